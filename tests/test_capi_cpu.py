"""CPU-side checks of the C ABI: the library loads without a GPU, exports every symbol declared in
include/sll_b200.h, fails loudly (no fallback) when a compute entry point is called without a device,
and the host-side layout / remap-plan logic reproduces sll_m_remapper's rules."""
import ctypes
import os
import re

import numpy as np
import pytest

import selalib_b200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for hdr in ("sll_b200.h", "sll_b200_sim6d_compat.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        syms |= set(re.findall(r"\b((?:sllb_|sim_bsl_vp_3d3v_cart_dd_slim_|sll_s_)[a-z0-9_]+)\s*\(", src))
    return sorted(syms)


def test_library_loads_and_exports_all_symbols():
    lib = sb.lib()
    syms = declared_symbols()
    assert len(syms) > 80 and "sim_bsl_vp_3d3v_cart_dd_slim_init" in syms and "sll_s_halt_collective" in syms
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert sb.device_count() == 0
    for ctor in (lambda: sb.Field([16, 16]), lambda: sb.Advector1dPeriodic(32, 0.0, 1.0),
                 lambda: sb.Poisson([32], [0.0], [1.0]), lambda: sb.Sim2d(32, 32, 0, 1, -1, 1, 0, 0.5, 1e-3, 0.1)):
        with pytest.raises(sb.SllbError) as e:
            ctor()
        assert e.value.code == 4 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "selalib_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in txt.replace("no CPU fallback", ""), fn


def test_factorize_and_process_grid():
    # sll_s_factorize_in_two_powers_of_two (sll_m_remapper.F90:6452-6476)
    assert [sb.factorize_in_two_powers_of_two(p) for p in (1, 2, 4, 8, 16, 32)] == \
        [(1, 1), (1, 2), (2, 2), (2, 4), (4, 4), (4, 8)]
    with pytest.raises(sb.SllbError):
        sb.factorize_in_two_powers_of_two(6)
    # sll_f_set_process_grid table (sll_m_decomposition.F90:2489-2531)
    table = {1: (1,) * 6, 2: (1, 1, 1, 1, 1, 2), 4: (1, 1, 1, 1, 2, 2), 8: (1, 1, 1, 2, 2, 2), 16: (1, 1, 2, 2, 2, 2),
             64: (2,) * 6, 128: (2, 2, 2, 2, 2, 4), 4096: (4,) * 6}
    for n, g in table.items():
        assert sb.set_process_grid(n) == g


def test_layout_boxes_split_rule():
    # 129 points over 2 -> 65 + 64 (odd branch, sll_m_remapper.F90:1921); rank = i + P1*(j + P2*(k + P3*l))
    b = sb.layout4d_boxes([129, 129, 33, 32], [2, 4, 1, 1], 8)
    assert b[0, 0].tolist() == [0, 64] and b[1, 0].tolist() == [65, 128]
    assert [b[2 * j, 1].tolist() for j in range(4)] == [[0, 32], [33, 64], [65, 96], [97, 128]]
    assert (b[:, 2] == [0, 32]).all() and (b[:, 3] == [0, 31]).all()
    # boxes tile the global array exactly once
    cover = np.zeros((129, 129), dtype=int)
    for r in range(8):
        cover[b[r, 0, 0]:b[r, 0, 1] + 1, b[r, 1, 0]:b[r, 1, 1] + 1] += 1
    assert (cover == 1).all()
    with pytest.raises(sb.SllbError):
        sb.layout4d_boxes([16] * 4, [3, 1, 1, 1], 3)


@pytest.mark.parametrize("nranks", [1, 2, 4, 8])
def test_remap_plan_moves_every_element_once(nranks):
    """test_remap_4d.F90:118-301: fill with the global linear index, remap, check every element."""
    g = [12, 9, 8, 10]
    f1, f2 = sb.factorize_in_two_powers_of_two(nranks)
    px, pv = [1, 1, f1, f2], [f1, f2, 1, 1]
    glob = np.arange(np.prod(g), dtype=float).reshape(g, order="F")
    bx = sb.layout4d_boxes(g, px, nranks)
    bv = sb.layout4d_boxes(g, pv, nranks)
    sl = lambda b: tuple(slice(b[d, 0], b[d, 1] + 1) for d in range(4))
    local_x = [glob[sl(bx[r])].copy(order="F") for r in range(nranks)]
    local_v = [np.full([bv[r, d, 1] - bv[r, d, 0] + 1 for d in range(4)], -1.0, order="F") for r in range(nranks)]
    for src in range(nranks):
        sboxes, _ = sb.remap4d_plan(g, px, pv, nranks, src)
        for dst in range(nranks):
            _, rboxes = sb.remap4d_plan(g, px, pv, nranks, dst)
            s, r = sboxes[dst], rboxes[src]
            assert (s == r).all()          # both sides agree on the intersection
            if (s[:, 0] > s[:, 1]).any():
                continue
            src_sl = tuple(slice(s[d, 0] - bx[src, d, 0], s[d, 1] + 1 - bx[src, d, 0]) for d in range(4))
            dst_sl = tuple(slice(s[d, 0] - bv[dst, d, 0], s[d, 1] + 1 - bv[dst, d, 0]) for d in range(4))
            local_v[dst][dst_sl] = local_x[src][src_sl]
    for r in range(nranks):
        assert np.array_equal(local_v[r], glob[sl(bv[r])])


def test_compat6d_check_reads_reference_format(tmp_path):
    """sll_s_check_diagnostics semantics on the golden file itself (host only): identical -> PASSED,
    perturbed beyond 5e-7 -> FAILED."""
    lib = sb.lib()
    gold = os.path.join(ROOT, "tests", "golden", "reffile_bsl_vp_3d3v_cart_dd.dat")
    assert lib.sllb_sim6d_compat_check(gold.encode(), gold.encode()) == 0
    rows = np.loadtxt(gold)
    rows[1, 3] += 1e-6
    bad = tmp_path / "bad.dat"
    np.savetxt(bad, rows)
    assert lib.sllb_sim6d_compat_check(gold.encode(), str(bad).encode()) != 0
    assert lib.sllb_sim6d_compat_check(gold.encode(), str(tmp_path / "missing.dat").encode()) != 0


def test_spline_dd_blocks_host_plan():
    """sllb_spline_dd_blocks / sllb_lagrange_dd_blocks (host only) against the literal restatement of
    make_blocks_spline (sll_m_advection_6d_spline_dd_slim.F90:202-287) and make_blocks_lagrange
    (sll_m_advection_6d_lagrange_dd_slim.F90:202-286)"""
    from oracle import orc
    rng = np.random.default_rng(20261017)
    for n in (16, 32, 33):
        v = -6.0 + 12.0 / n * np.arange(n)
        for fac in (0.04, 0.2, 0.45, -0.3, 0.25):      # 0.25: displacements that are exactly integer
            disp = -v * fac
            shift, alpha, nb = sb.spline_dd_blocks(disp)
            oshift, oalpha, onb = orc.make_blocks_spline(disp)
            assert np.array_equal(shift, oshift) and nb == onb
            assert np.array_equal(alpha, oalpha)
    # monotonic random displacements without zeros
    disp = np.sort(rng.uniform(-2.7, 1.9, 40))
    for d in (disp, disp[::-1].copy()):
        shift, _, nb = sb.spline_dd_blocks(d)
        oshift, _, onb = orc.make_blocks_spline(d)
        assert np.array_equal(shift, oshift) and nb == onb
    # centred Lagrange: halo widths stencil/2 - box - 1, stencil/2 + box per block (:239-240,257-258)
    v = -6.0 + 12.0 / 32 * np.arange(32)
    box, nb, hw = sb.lagrange_dd_blocks(-v * 0.3, 6)
    assert nb == 4 and np.array_equal(hw, [[1, 4], [2, 3], [3, 2], [4, 1]])
    assert box[16] == sb.SHIFT_SKIP and box[0] == 1 and box[-1] == -2
    with pytest.raises(sb.SllbError):
        sb.lagrange_dd_blocks(-v * 0.3, 2)          # displacement leaves [-stencil/2, stencil/2)


def test_splitting_tables_match_independent_transcription():
    """sllb_splitting_coeff (half-palindromes + dt polynomials) against the oracle's literal transcription of
    sll_m_time_splitting_coeff.F90:157-594, every split_case, several dt; plus structural checks."""
    from oracle import orc
    for k, name in enumerate(orc.SPLIT_CASES):
        assert sb.splitting_case(name) == k
        for dt in (0.0, 0.05, 0.1, 0.25):
            steps, bt, dv, nb = sb.splitting_coeff(name, dt)
            osteps, obt, odv = orc.splitting_coeff(name, dt)
            assert bt == obt and dv == odv and steps.size == osteps.size
            assert np.abs(steps - osteps).max() <= 4e-16, name
        # consistency of the schemes at dt = 0: T weights and V weights each sum to one
        steps, bt, dv, nb = sb.splitting_coeff(name, 0.0)
        T, V, idx, isT = [], [], 0, bt
        for _ in range(nb):
            if isT:
                T.append(steps[idx]); idx += 1
            else:
                V.append(steps[idx]); idx += dv
            isT = not isT
        assert idx == steps.size
        assert abs(sum(T) - 1) < 1e-14 and abs(sum(V) - 1) < 1e-14, name
    with pytest.raises(sb.SllbError):
        sb.splitting_case("SLL_NOT_A_CASE")
    # finite-difference weights of compute_jacobian (sll_s_compute_w_hermite)
    assert np.allclose(sb.compute_w_hermite(-2, 2), [1 / 12, -2 / 3, 0, 2 / 3, -1 / 12], atol=1e-16)
    for r, s in ((-1, 1), (-2, 2), (-3, 3), (-2, 3)):
        assert np.abs(sb.compute_w_hermite(r, s) - orc.compute_w_hermite(r, s)).max() < 1e-15
        k = np.arange(r, s + 1)
        w = sb.compute_w_hermite(r, s)
        assert abs(w.sum()) < 1e-15 and abs((w * k).sum() - 1) < 1e-14      # exact on constants and on x


def test_fortran_g20_12_formatter():
    """the '(13g20.12)' rows of thdiag.dat (sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:998-1010): Fortran G editing,
    checked on values as they appear in the reference's own vpsim4d_cartesian_ref.dat columns"""
    cases = {0.0: "   0.00000000000    ", 1.00003947842: "   1.00003947842    ", 157.913693328: "   157.913693328    ",
             0.789568343347E-04: "  0.789568343347E-04", -0.5: " -0.500000000000    ", 0.1: "  0.100000000000    ",
             123456789012.4: "   123456789012.    ", 1e12: "  0.100000000000E+13", -3.25e-7: " -0.325000000000E-06",
             0.0999999999999: "  0.999999999999E-01"}
    for x, s in cases.items():
        assert sb.format_g20_12(x) == s, (x, sb.format_g20_12(x))
        assert len(sb.format_g20_12(x)) == 20


def test_fortran_g25_15_formatter():
    """the '(8g25.15)' / '(1g25.15)' cells of the 1D1V thdiag.dat (sll_m_sim_bsl_vp_1d1v_cart.F90:1778-1801)"""
    assert sb.format_g(0.1, 25, 15) == "    0.100000000000000    "
    assert sb.format_g(12.5663706012, 25, 15) == "     12.5663706012000    "
    assert sb.format_g(-57.4593058704, 25, 15) == "    -57.4593058704000    "
    assert sb.format_g(0.124099195637E-04, 25, 15) == "    0.124099195637000E-04"
    assert sb.format_g(0.0, 25, 15) == "     0.00000000000000    "
    for x in (1.0, 136.685004002, -3.3e-9, 7.25e17, 0.999999999999999999):
        c = sb.format_g(x, 25, 15)
        assert len(c) == 25 and abs(float(c) / x - 1) < 1e-14


def test_namelist_front_end_host_logic(tmp_path):
    """sllb_sim4d_create_from_namelist: parsing, defaults and the reference's error messages happen on the host, before a
    device is needed (sll_m_sim_bsl_vp_2d2v_cart_poisson_serial.F90:300-624)"""
    import ctypes as C
    lib = sb.lib()

    def create(text, name="in"):
        path = tmp_path / (name + ".nml")
        path.write_text(text)
        S = C.c_void_p()
        nit, fdt = C.c_int(-1), C.c_int(-1)
        rc = lib.sllb_sim4d_create_from_namelist(str(tmp_path / name).encode(), None, C.byref(S), C.byref(nit), C.byref(fdt))
        if rc == 0:
            lib.sllb_sim4d_destroy(S)
        return rc, sb.last_error(), nit.value, fdt.value

    ok = "&geometry\n num_cells_x1 = 16\n/\n&time_iterations\n dt = 0.1\n number_iterations = 7\n freq_diag_time = 2\n/\n"
    rc, msg, nit, fdt = create(ok)
    assert rc in (0, sb.ERR_NO_DEVICE), msg          # parsed; only the device may be missing
    if rc == 0:
        assert (nit, fdt) == (7, 2)
    rc, msg, _, _ = create(ok + "&advector\n advector_x3 = \"SLL_FOO\"\n/\n", "bad_adv")
    assert rc == sb.ERR_UNSUPPORTED and "advector in x3" in msg and "not implemented" in msg
    rc, msg, _, _ = create("&time_iterations\n split_case = \"SLL_NOPE\"\n/\n", "bad_split")
    assert rc == sb.ERR_INVALID and "split_case not defined" in msg
    rc, msg, _, _ = create("&geometry\n mesh_case_x3 = \"SLL_LANDAU_MESH\"\n/\n", "bad_mesh")
    assert rc == sb.ERR_UNSUPPORTED and "mesh_case_x3" in msg
    rc, msg, _, _ = create("&initial_function\n initial_function_case = \"SLL_BEAM\"\n/\n", "bad_init")
    assert rc == sb.ERR_UNSUPPORTED
    S = C.c_void_p()
    rc = lib.sllb_sim4d_create_from_namelist(str(tmp_path / "does_not_exist").encode(), None, C.byref(S), None, None)
    assert rc == sb.ERR_INVALID and "failed to open file" in sb.last_error()


def test_public_headers_compile_as_c_and_cpp(tmp_path):
    """include/*.h are the drop-in boundary: plain C (what a cgo / iso_c_binding / ctypes binding reads) and C++"""
    import subprocess
    src = tmp_path / "h.c"
    src.write_text('#include "sll_b200.h"\n#include "sll_b200_sim6d_compat.h"\nint main(void) { return SLLB_SHIFT_SKIP == INT32_MIN ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, "-x", "c++", str(src)])


def test_pipelined_exchange_chunks_tile_the_lines():
    """sllb_dd6d_chunk_boxes: the pieces of a pipelined split-axis pass cover every line exactly once, inner ranges are
    multiples of 32 lines (TMA rows), and shapes that cannot be cut stay in one piece"""
    import ctypes as C
    lib = sb.lib()
    for outer, inner, nch in ((256, 1024, 4), (16, 4096, 4), (1, 32 * 32 * 32 * 16 * 16, 4), (1, 8388608, 8), (3, 100, 4), (1, 96, 4),
                              (1, 50, 4), (5, 64, 16)):
        buf = (C.c_longlong * 64)(); nb = C.c_int(0)
        assert lib.sllb_dd6d_chunk_boxes(C.c_longlong(outer), C.c_longlong(inner), C.c_int(nch), buf, C.byref(nb)) == 0
        boxes = np.array(buf[:4 * nb.value]).reshape(-1, 4)
        cover = np.zeros((outer, min(inner, 4096)), dtype=int)
        for o0, oc, i0, ic in boxes:
            assert oc >= 1 and ic >= 1 and o0 >= 0 and i0 >= 0 and o0 + oc <= outer and i0 + ic <= inner
            if ic != inner:
                assert i0 % 32 == 0 and ic % 32 == 0
            cover[o0:o0 + oc, min(i0, cover.shape[1]):min(i0 + ic, cover.shape[1])] += 1
        assert (cover == 1).all()
        assert sum(oc * ic for _, oc, _, ic in boxes) == outer * inner
        assert 1 <= nb.value <= nch


def test_new_entry_points_fail_loudly_without_a_device():
    """no CPU fallback behind the (f) entry points either: without a CUDA device they return SLLB_ERR_NO_DEVICE"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    for ctor in (lambda: sb.Interpolator1d(sb.INTERP_CUBIC_SPLINE, 33, -6.0, 6.0, bc=sb.BC_HERMITE),
                 lambda: sb.Sim4d([16, 16, 32, 32], [0, 0, -6, -6], [12.5, 12.5, 6, 6], 0.5, 0.5, 1e-3, 0.1, split="SLL_ORDER6VPOT_VTV"),
                 lambda: sb.Sim6d([16] * 6, 6.0, [12.5] * 3, 3, 3, 0.01, 0.01, [0.5] * 3, advector=sb.ADVECTOR_SPLINE),
                 lambda: sb.Dd6d(None, [16] * 6)):
        with pytest.raises(sb.SllbError) as e:
            ctor()
        assert e.value.code == sb.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_tuning_setters_reject_values_outside_their_range():
    """the kernel-variant switches validate their arguments on the host (no device needed) and keep the defaults"""
    for bad in ((2, 1), (-2, 1), (0, 2), (0, -1)):
        with pytest.raises(sb.SllbError) as e:
            sb.set_plane_variant(*bad)
        assert e.value.code == sb.ERR_INVALID
    with pytest.raises(sb.SllbError):
        sb.set_plane_kernel(True, 8)
    sb.set_plane_variant(-1, 1)     # the defaults: accepted
    sb.set_plane_kernel(True, 0)


def _build_c_driver(tmp_path):
    import subprocess
    libdir = os.path.join(ROOT, "selalib_b200", "lib")
    exe = str(tmp_path / "sim_bsl_vp_2d2v_cart_poisson_serial_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "sim_bsl_vp_2d2v_cart_poisson_serial_b200.c"), "-o", exe,
                           "-L" + libdir, "-lsllb200", "-Wl,-rpath," + libdir])
    return exe


def test_c_driver_builds_and_fails_loudly_without_a_device(tmp_path):
    """tests/c/sim_bsl_vp_2d2v_cart_poisson_serial_b200.c: a plain-C host of the drop-in boundary links against the
    library; without a GPU it stops with the library's message instead of computing anything"""
    import subprocess
    import torch
    exe = _build_c_driver(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("a GPU is present (covered by tests/test_gpu_splitting.py)")
    out = subprocess.run([exe, str(tmp_path / "whatever")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 1 and "no usable CUDA device" in out.stderr
