#!/usr/bin/env python
"""bench.py -- headline benchmark of the split semi-Lagrangian advection path.

Metric (BASELINE.json): 4D phase-space point-updates/s per advection pass, on the 2D2V Landau-damping
128^4 fp64 configuration (cubic-spline BSL, Strang VTV, FFT Poisson between the splitting stages).
One "step" = one Strang time step = 6 advection passes over all of f (V/2: x3,x4; T: x1,x2; V/2: x3,x4)
plus 2 charge-density reductions and 2 Poisson solves; value = 6 * N^4 * steps / time.
Strong scaling: the same 128^4 problem on N GPUs (x <-> v remaps over NCCL).

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (CPU restatement of the reference path)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "4D phase-space point-updates/s per advection pass"
UNIT = "point-updates/s"
NSIDE = int(os.environ.get("SLLB_BENCH_N", "128"))
PASSES_PER_STEP = 6
KERNEL_ALONE_PAUSE_S = 0.5   # idle time before each per-kernel timing block of the roofline section
XMIN = [0.0, 0.0, -6.0, -6.0]
XMAX = [4 * np.pi, 4 * np.pi, 6.0, 6.0]
WORKLOAD = (f"2D2V Landau damping {NSIDE}^4 fp64, periodic cubic-spline BSL on all four axes, Strang VTV, "
            "trapezoid rho + 2D FFT Poisson (sim_bsl_vp_2d2v_cart_poisson_serial semantics), dt=0.1, eps=1e-3, k=(0.5,0.5)")


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi sampler running while the GPU sections execute (the recipe's clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        # median over the samples taken under load (upper half by power draw)
        if sm:
            order = np.argsort(power)
            hot = [sm[i] for i in order[len(order) // 2:]]
            med = statistics.median(hot)
        else:
            med = None
        return {"sm_mhz": med, "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": "warm-up + timed steps + per-kernel timing + e2e"}


def cpu_sample_step(orc, f, n, frac, method="spline"):
    """One Strang step's 6 advection passes (+ 2 Poisson solves) of the CPU restatement over 1/frac of the
    lines of the n^4 field, full-length lines.  Returns point-updates done."""
    v = XMIN[2] + (XMAX[2] - XMIN[2]) / n * np.arange(n)
    dx = (XMAX[0] - XMIN[0]) / n
    dv = (XMAX[2] - XMIN[2]) / n
    dt = 0.1
    E = 1e-3 * np.sin(np.arange(n * n) * 0.01)
    rho = np.zeros((n, n), order="F")
    done = 0

    def vstage(step):
        nonlocal done
        orc.poisson_2d(rho, n, n, XMIN[0], XMAX[0], XMIN[1], XMAX[1])
        done += orc.advect_axis_sub(f, 2, method, 4, E * (-step * dt / dv), (1, 1, 0, 1, n * n, 1), frac)
        done += orc.advect_axis_sub(f, 3, method, 4, E * (-step * dt / dv), (1, 1, 0, 1, n * n, 1), frac)

    vstage(0.5)
    done += orc.advect_axis_sub(f, 0, method, 4, v * (-dt / dx), (n, n, 1, 1, 1, 0), frac)
    done += orc.advect_axis_sub(f, 1, method, 4, v * (-dt / dx), (n, n, 1, 1, 1, 0), frac)
    vstage(0.5)
    return done


def cpu_field(n):
    rng = np.random.default_rng(20261017)
    f = np.empty((n, n, n, n), order="F")
    flat = f.reshape(-1, order="F")
    chunk = 1 << 24
    for i in range(0, flat.size, chunk):
        flat[i:i + chunk] = rng.random(min(chunk, flat.size - i))
    return f


def run_reference(args, rank):
    """CPU arm: the oracle port of the reference path (the reference itself is Fortran and cannot be built
    in this image), all host threads, on bounded samples of the same workload."""
    if rank != 0:
        return
    from oracle import orc
    # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for the host's cores explicitly
    cores = orc.set_num_threads(env_int("SLLB_REF_THREADS", 0) or orc.host_cores())
    n = NSIDE
    f = cpu_field(n)
    # size the per-step sample so that the whole run ends within a few minutes whatever K is
    t0 = time.perf_counter()
    d0 = cpu_sample_step(orc, f, n, 64)
    thr = d0 / (time.perf_counter() - t0)
    budget_s = float(os.environ.get("SLLB_REF_BUDGET_S", "90"))
    full = PASSES_PER_STEP * float(n) ** 4
    frac = env_int("SLLB_REF_FRAC", 0) or int(min(4096, max(4, np.ceil((args.steps + args.warmup) * full / (thr * budget_s)))))
    for _ in range(args.warmup):
        cpu_sample_step(orc, f, n, frac)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        done += cpu_sample_step(orc, f, n, frac)
    dtm = time.perf_counter() - t0
    value = done / dtm
    sample = (f"per step: the 6 advection passes of one Strang step over 1/{frac} of the lines (full {n}-point lines, all four "
              f"axes) of the {n}^4 field + 2 Poisson solves; direct cubic-spline algorithm (sll_m_cubic_splines fast path), "
              f"line copy-in/out, OpenMP static over lines on {cores} threads; rho reduction excluded")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dtm / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "arm": "C restatement of the reference CPU path (oracle port), host cores"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def c5_block(sb, torch, dist, comm, rank, world, peak):
    """C5 (BASELINE config 5): 3D3V Landau damping 32^6 fp64, domain-decomposed 7-point Lagrange advection with halo
    exchange over NVLink (sim_bsl_vp_3d3v_cart_dd_slim semantics), timed like the headline: CUDA events, max over ranks."""
    n = env_int("SLLB_C5_N", 32)
    steps, warmup = env_int("SLLB_C5_STEPS", 5), 3

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    args6 = ([n] * 6, 6.0, [4 * np.pi] * 3, 7, 7, 0.01, 0.01, [0.5] * 3)
    S = sb.Sim6d(*args6, comm=comm, time_in_phase=False)
    lay = S.layout()
    r_w = S.run(warmup)
    S.halo_ms()
    barrier()
    sb.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r_t = S.run(steps)
    e1.record()
    barrier()
    ms = maxr(e0.elapsed_time(e1))
    launches = sb.launch_count()
    rows = np.vstack([r_w, r_t])
    reps = 3
    S.halo_ms()
    barrier()
    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    x0.record()
    for _ in range(reps):
        S.advect_x()
    x1.record()
    barrier()
    x_ms = maxr(x0.elapsed_time(x1)) / (3 * reps)
    S.halo_ms()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    v0.record()
    for _ in range(reps):
        S.advect_v(0.01)
    v1.record()
    barrier()
    v_ms = maxr(v0.elapsed_time(v1)) / (3 * reps)
    # exchange time per split pass: opt-in (one host synchronisation per pass), measured in a loop of its own
    sb.dd6d_set_exchange_timing(True)
    S.halo_ms()
    for _ in range(reps):
        S.advect_v(0.01)
    nsplit = sum(1 for p in lay["procs"][3:] if p > 1)
    halo_ms = maxr(S.halo_ms()) / max(1, nsplit * reps)
    sb.dd6d_set_exchange_timing(False)
    barrier()
    S.destroy()
    local_pts = float(np.prod(lay["nw"]))
    npts = float(n) ** 6
    # cross-N exactness: rank 0 repeats the same steps on ONE GPU (8.6 GB) and compares the 14-column rows
    rows_vs_1gpu = None
    rows_vs_1gpu_cols = None
    if world > 1 and rank == 0 and not os.environ.get("SLLB_SKIP_1GPU_CHECK"):
        S1 = sb.Sim6d(*args6, time_in_phase=False)
        r1 = S1.run(warmup + steps)
        S1.destroy()
        # the 14 columns of the reference's .dat file; the ones that are rounding noise around zero (momenta, ...) are
        # measured against the largest column instead of against themselves
        # against the largest column: the file holds integrals of very different size (mass ~1, field energy ~1e-9, momenta
        # that are rounding noise around zero), and what a misplaced halo plane would change is f itself
        rows_vs_1gpu = float(np.abs(rows - r1).max() / np.abs(r1[:, 1:]).max())
        rows_vs_1gpu_cols = (np.abs(rows - r1).max(axis=0) / np.maximum(np.abs(r1).max(axis=0), 1e-300)).tolist()
    barrier()
    hw = 3
    halo_bytes = 2 * hw * local_pts / lay["nw"][5] * 8 if world > 1 else 0
    return {"workload": f"3D3V Landau damping {n}^6 fp64, Lagrange fixed 7-point in all six directions, dt=0.01, landau_prod "
                        "alpha=0.01 k=0.5, L=4pi, v_max=6 (sim_bsl_vp_3d3v_cart_dd_slim semantics), strong scaling",
            "metric": "6D phase-space point-updates/s per advection pass", "value": 6 * npts * steps / (ms * 1e-3),
            "ms_per_step": ms / steps, "steps": steps, "warmup": warmup, "process_grid": lay["procs"], "local_block": lay["nw"],
            "x_pass_ms": x_ms, "x_pass_gbs": 16 * local_pts / (x_ms * 1e-3) / 1e9,
            "x_pass_frac_of_measured_hbm": 16 * local_pts / (x_ms * 1e-3) / 1e9 / peak,
            "v_pass_ms": v_ms, "v_pass_gbs": 16 * local_pts / (v_ms * 1e-3) / 1e9,
            "halo_ms_per_split_pass": halo_ms, "halo_bytes_sent_per_split_pass": halo_bytes,
            "halo_gbs_per_direction": (halo_bytes / 2) / (halo_ms * 1e-3) / 1e9 if halo_ms > 0 else None,
            "frac_of_aggregate_hbm_roofline_whole_step": 16.0 * 6 * npts * steps / (ms * 1e-3) / 1e9 / (peak * world),
            "gpu_launches": int(launches),
            "check": {"mass": float(rows[-1, 1]), "l2": float(rows[-1, 2]), "rows_vs_1gpu_rel": rows_vs_1gpu,
                      "rows_vs_1gpu_rel_per_column": rows_vs_1gpu_cols,
                      "what": "the 14 columns of the reference's .dat rows after warmup + steps steps, N GPUs against 1 GPU: largest "
                              "difference relative to the largest column, and column by column (columns that are rounding noise "
                              "around zero compare noise with noise)"}}


def run_ours(args, rank, world, local_rank):
    import ctypes as C

    import torch
    import selalib_b200 as sb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; selalib_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    sb.init(local_rank)
    comm = None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.tensor(list(sb.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = sb.Comm(bytes(idt.cpu().tolist()), world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = NSIDE
    if os.environ.get("SLLB_REMAP_ROTATION") is not None:
        sb.set_remap_rotation(env_int("SLLB_REMAP_ROTATION", 1))
    sampler = ClockSampler(local_rank) if rank == 0 else None
    sim_args = ([n] * 4, XMIN, XMAX, 0.5, 0.5, 1e-3, 0.1)
    sim_kw = dict(split=0, method=sb.METHOD_SPLINE, order=4)
    S = sb.Sim4d(*sim_args, comm=comm, **sim_kw)
    npts = float(n) ** 4

    # ---- kernel-resident timing: inputs already in HBM; the diagnostics row of every step (field energy, mass, L1, L2,
    # kinetic energy -- the reference computes them every step) is inside the timed region ---------------------------
    sb.set_phase_timers(False)
    S.run(args.warmup, diagnostics=True)
    barrier()
    sb.launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rows_t = S.run(args.steps, diagnostics=True)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = sb.launch_count()
    value = PASSES_PER_STEP * npts * args.steps / (ms * 1e-3)
    # cross-N exactness signal: the state after exactly warmup + steps steps from the initial data, before anything else
    # touches f -- identical (to rounding) whatever the number of GPUs
    csum = S.checksum()
    check = {"steps_from_initial_data": args.warmup + args.steps, "mass": float(rows_t[-1, 3]), "field_energy": float(rows_t[-1, 1]),
             "l2": float(rows_t[-1, 5]), "checksum_wf": float(csum[0]), "checksum_wf2": float(csum[1]),
             "what": "after the timed region; checksum = sum w f and sum w f^2 with w a function of the GLOBAL index of every point"}
    # the same region without the diagnostics rows, and the per-phase breakdown (pooled CUDA events, opt-in)
    S.run(min(args.warmup, 3), diagnostics=False)   # its own recorded step is captured here, not inside the region
    barrier()
    e0.record()
    S.run(args.steps, diagnostics=False)
    e1.record()
    barrier()
    ms_nodiag = max_over_ranks(e0.elapsed_time(e1))
    sb.set_phase_timers(True)
    ph_steps = min(args.steps, 10)
    S.run(ph_steps, diagnostics=True)
    phase = (S.phase_ms8() / ph_steps).tolist()
    sb.set_phase_timers(False)

    # ---- per-kernel timing for the roofline (CUDA events on the launch stream, local field) ------
    F = S.field()
    ext = F.extents
    local_pts = float(np.prod(ext))
    reps = 20
    dispv = torch.linspace(-2.3, 2.3, max(ext), dtype=torch.float64, device="cuda")
    Ef = (1e-2 * torch.sin(torch.arange(ext[0] * ext[1], dtype=torch.float64, device="cuda"))).contiguous()
    kernel_ms = {}
    axes = [0, 1] + ([2, 3] if world == 1 else [])
    for axis in axes:
        if axis < 2:
            dsel = (ext[1] if axis == 0 else 1, ext[axis + 2], 1, 1, 1, 0)
            call = lambda a=axis, d=dsel: F.advect_axis(a, sb.METHOD_SPLINE, 4, dispv.data_ptr(), 1.0, d, on_device=True)
        else:
            dsel = (1, 1, 0, 1, ext[0] * ext[1], 1)
            call = lambda a=axis, d=dsel: F.advect_axis(a, sb.METHOD_SPLINE, 4, Ef.data_ptr(), 1.0, d, on_device=True)
        # every kernel is timed ALONE: the board reaches its power cap ~0.1 s into back-to-back launches and then holds
        # the SM clock near 1.6 GHz (profiles/r02_plane_sustained.log), so a pause lets each block of 23 launches
        # (15 ms) start from an idle board -- the state MEASURED_PEAKS.json's burst copy bandwidth was taken in
        torch.cuda.synchronize()
        time.sleep(KERNEL_ALONE_PAUSE_S)
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            call()
        a1.record()
        torch.cuda.synchronize()
        kernel_ms[axis] = a0.elapsed_time(a1) / reps
    # the T stage runs as ONE kernel on a single GPU (x1 pass + x2 pass + charge density, K1c): time it too
    plane_ms = None
    if world == 1:
        from selalib_b200.capi import DispT, dp as _dp, vp as _vp
        import ctypes as _C

        def _disp(dsel):
            d = DispT()
            d.values = _C.cast(_vp(dispv.data_ptr()), _dp); d.nvalues = 0; d.values_on_device = 1; d.scale = 1.0
            d.odiv, d.omod, d.ostr, d.idiv, d.imod, d.istr = dsel
            return d
        pd0, pd1 = _disp((ext[1], ext[2], 1, 1, 1, 0)), _disp((ext[2], ext[3], 1, 1, 1, 0))
        rho_t = torch.empty(ext[0] * ext[1], dtype=torch.float64, device="cuda")

        def plane_call():
            assert sb.lib().sllb_advect_plane(F.h, 0, 4, _C.byref(pd0), _C.byref(pd1), _C.c_double(1.0),
                                              _C.cast(_vp(rho_t.data_ptr()), _dp)) == 0, sb.last_error()
        torch.cuda.synchronize()
        time.sleep(KERNEL_ALONE_PAUSE_S)
        for _ in range(3):
            plane_call()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(reps):
            plane_call()
        a1.record()
        torch.cuda.synchronize()
        plane_ms = a0.elapsed_time(a1) / reps
    # dominant kernel: the strided spline pass.  One GPU: the x3 and x4 passes of both V stages (4 of the 6 passes of a
    # step), timed on axes 2 and 3.  Several GPUs: the field handed out here is the x-sequential box, so the same kernel is
    # timed on its strided x2 axis (the x3/x4 passes run it in the other layout, x4 with the remap stores fused in).
    if world == 1:
        strided, timed_on = [kernel_ms[2], kernel_ms[3]], "x3 and x4 passes (4 of the 6 passes of a step; the other two are the single k_spline_plane_r launch of the T stage)"
    else:
        strided, timed_on = [kernel_ms[1]], "timed in place on the strided x2 axis of the local x-sequential box (in the step it runs the x3 pass and, with the remap stores fused in, the x4 pass)"
    t_dom = sum(strided) / len(strided)
    peak, peak_src = measured_peak()
    achieved = 16.0 * local_pts / (t_dom * 1e-3) / 1e9
    traffic, traffic_src, plane_traffic = None, None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            # the plain variant of the strided spline pass: <P = 4, REMAP = 0, DIAG = 0[, compile-time line length]>
            names = [k for k in tj["kernels"] if k.startswith("k_spline_strided_split<4") and
                     all(t.strip() == "0" for t in k[k.index("<") + 1:k.rindex(">")].split(",")[1:3])]
            names.sort(key=lambda k: -tj["kernels"][k]["launches"])
            traffic = tj["kernels"][names[0]]["dram_bytes_per_launch"] * local_pts / float(128) ** 4
            traffic_src = tj.get("source")
            t_plane = [k for k in tj["kernels"] if k.startswith("k_spline_plane_r<1")]
            plane_traffic = tj["kernels"][t_plane[0]]["dram_bytes_per_launch"] * local_pts / float(128) ** 4 if t_plane else None
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": f"k_spline_strided_split<4>: {timed_on}",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": 16.0 * local_pts,
                "ms_per_launch": t_dom, "ms_per_launch_by_axis": {f"x{a + 1}": kernel_ms[a] for a in axes},
                "gbs_by_axis": {f"x{a + 1}": 16.0 * local_pts / (kernel_ms[a] * 1e-3) / 1e9 for a in axes},
                "t_stage_plane_kernel": None if plane_ms is None else {
                    "kernel": "k_spline_plane_r<rho> (x1 pass + x2 pass + charge density in one sweep) + the sum of its per-CTA partial densities", "ms_per_launch": plane_ms,
                    "algorithmic_bytes_per_launch": 32.0 * local_pts, "achieved_gbs_at_16B_per_point_per_pass": 32.0 * local_pts / (plane_ms * 1e-3) / 1e9,
                    "hbm_bytes_moved_per_launch": 16.0 * local_pts, "hbm_gbs": 16.0 * local_pts / (plane_ms * 1e-3) / 1e9,
                    "frac_of_measured_hbm": 16.0 * local_pts / (plane_ms * 1e-3) / 1e9 / peak, "traffic": plane_traffic},
                # 6 passes at 16 B per point over the step time (above 1 on one GPU: the T stage moves f once for two passes),
                # and the same with the 5 sweeps over f a step really makes on one GPU (x1+x2 fused, x3, x4, x3, x4)
                "whole_step_frac_of_aggregate_hbm_roofline": 16.0 * value / 1e9 / (peak * world),
                "whole_step_frac_of_hbm_roofline_5_sweeps": (16.0 * value * 5.0 / 6.0 / 1e9 / peak) if world == 1 else None,
                "timing": f"CUDA events on the launch stream, {reps} launches after 3 warm-ups, every kernel alone after a "
                          f"{KERNEL_ALONE_PAUSE_S} s pause (the power cap pulls the SM clock ~0.1 s into sustained load), field {ext} "
                          f"({local_pts * 8 / 1e9:.2f} GB > L2)"}

    # ---- end to end through the C ABI with HOST buffers ------------------------------------------
    host = torch.empty(int(local_pts), dtype=torch.float64).pin_memory()
    lib = sb.lib()
    hp = C.cast(C.c_void_p(host.data_ptr()), C.POINTER(C.c_double))
    assert lib.sllb_field_download(F.h, hp, None) == 0
    F = S.field()
    e2e_steps = max(1, min(args.steps, env_int("SLLB_E2E_STEPS", 10)))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        assert lib.sllb_field_upload(F.h, hp, None) == 0          # H2D of the step's input state (pinned)
        S.run(1, diagnostics=True)                                 # one Strang step + diagnostics row (D2H)
        F = S.field()                                              # x-sequential layout (remap back when N > 1)
        assert lib.sllb_field_download(F.h, hp, None) == 0         # D2H of the step's result
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = PASSES_PER_STEP * npts * e2e_steps / e2e_s
    # ensemble streaming (single GPU): the same per-step traffic, but the upload of the next state and the download of the
    # previous one overlap the current state's step (sllb_sim4d_stream_step; PCIe is full duplex).  Fill and drain
    # calls are inside the timed region: every counted state is uploaded, stepped and downloaded within it.
    stream_value = None
    stream_note = None
    if world == 1 and not os.environ.get("SLLB_SKIP_STREAM"):
        try:
            host_out = torch.empty(int(local_pts), dtype=torch.float64).pin_memory()
            members = max(e2e_steps, env_int("SLLB_STREAM_MEMBERS", 12))
            S.stream_step(host.data_ptr(), None); S.stream_step(host.data_ptr(), None); S.stream_step(None, host_out.data_ptr())  # warm-up
            S.stream_step(None, None)
            barrier()
            t0 = time.perf_counter()
            for k in range(members + 2):
                S.stream_step(host.data_ptr() if k < members else None, host_out.data_ptr() if k >= 2 else None)
            barrier()
            stream_s = time.perf_counter() - t0
            stream_value = PASSES_PER_STEP * npts * members / stream_s
            del host_out
        except (RuntimeError, sb.SllbError) as exc:   # e.g. not enough pinned host or device memory for the two extra copies
            stream_value = None
            stream_note = f"ensemble streaming not measured ({exc}); value is the serial figure"
    e2e = {"value": stream_value if stream_value is not None else e2e_value, "unit": UNIT,
           "h2d_bytes_per_step": int(local_pts * 8), "d2h_bytes_per_step": int(local_pts * 8 + (0 if stream_value is not None else 48)),
           "steps": e2e_steps,
           "what": ("ensemble streaming through sllb_sim4d_stream_step: per step one state goes pinned host -> HBM, one state is "
                    "advanced by a full Strang step, one state comes back HBM -> pinned host; the two copies overlap the step "
                    "(three device copies of f rotate), fill + drain calls included in the timed region"
                    if stream_value is not None else
                    "per step: sllb_field_upload (pinned host f -> HBM) + sllb_sim4d_run(1 step, diagnostics) + "
                    "sllb_field_download (HBM -> pinned host f); per-rank local box"),
           "serial_value": e2e_value, "note": stream_note,
           "serial_what": "per step, one after the other: sllb_field_upload (pinned host f -> HBM) + sllb_sim4d_run(1 step, "
                          "diagnostics) + sllb_field_download (HBM -> pinned host f); per-rank local box",
           "resident_value": value,
           "resident_what": "the headline value: sllb_sim4d_run with f resident in HBM, one 6-double diagnostics row per step"}
    S.destroy()

    # ---- cross-N exactness of the headline run: rank 0 repeats the same steps on ONE GPU and compares -------------
    if world > 1 and not os.environ.get("SLLB_SKIP_1GPU_CHECK"):
        if rank == 0:
            S1 = sb.Sim4d(*sim_args, comm=None, **sim_kw)
            S1.run(args.warmup, diagnostics=True)
            r1 = S1.run(args.steps, diagnostics=True)
            c1 = S1.checksum()
            S1.destroy()
            ref = {"mass": float(r1[-1, 3]), "field_energy": float(r1[-1, 1]), "l2": float(r1[-1, 5]),
                   "checksum_wf": float(c1[0]), "checksum_wf2": float(c1[1])}
            check["vs_1gpu_rel"] = {k: abs(check[k] - v) / max(abs(v), 1e-300) for k, v in ref.items()}
            check["vs_1gpu_max_rel"] = max(check["vs_1gpu_rel"].values())
        barrier()

    # ---- C5: the 3D3V domain-decomposed Lagrange path on the same GPUs (BASELINE config 5) -------
    c5 = None
    if not os.environ.get("SLLB_SKIP_C5"):
        try:
            c5 = c5_block(sb, torch, dist, comm, rank, world, peak)
        except (RuntimeError, sb.SllbError) as exc:
            c5 = {"error": str(exc)}
    clocks = sampler.stop() if sampler else None

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not os.environ.get("SLLB_SKIP_CPU"):
        from oracle import orc
        orc.set_num_threads(orc.host_cores())
        frac = env_int("SLLB_REF_FRAC", 4)
        fc = cpu_field(n)
        cpu_sample_step(orc, fc, n, frac * 8)
        t0 = time.perf_counter()
        done = cpu_sample_step(orc, fc, n, frac)
        t_direct = time.perf_counter() - t0
        t0 = time.perf_counter()
        done_fft = cpu_sample_step(orc, fc, n, frac * 4, method="fft_spline")
        t_fft = time.perf_counter() - t0
        cpu_baseline = {"value": done / t_direct, "unit": UNIT, "cores": orc.num_threads(), "kind": "port",
                        "sample": f"6 advection passes of one Strang step over 1/{frac} of the lines of the {n}^4 field (full-length "
                                  "lines, all four axes) + 2 Poisson solves; C restatement of the reference CPU path, direct "
                                  "cubic-spline algorithm, OpenMP over lines",
                        "reference_algorithm_value": done_fft / t_fft,
                        "reference_algorithm": f"same passes over 1/{frac * 4} of the lines with the FFT-diagonalised periodic spline "
                                               "(sll_s_periodic_interp, the sims' default SLL_SPLINES advector)"}
        del fc

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "passes_per_step": PASSES_PER_STEP, "points": int(npts),
                           "timed_region": "sllb_sim4d_run(steps, diagnostics every step): 6 advection passes + 2 charge densities + 2 Poisson solves + the diagnostics row per step",
                           "l2": f"inputs larger than L2: f is {npts * 8 / 1e9:.2f} GB ({local_pts * 8 / 1e9:.2f} GB per GPU) vs 126 MB L2",
                           "parallelism": "single GPU" if world == 1 else f"{world} GPUs, x<->v remap fused into the last advection pass of each stage (peer stores over NVLink, CUDA IPC); 2 remaps per Strang step",
                           "staging": "TMA bulk copies (cp.async.bulk, UBLKCP): 256 B rows of 32-line tiles (V stage), whole 128x128 planes (T stage)"},
                "ms_per_step_without_diagnostics": ms_nodiag / args.steps,
                "value_without_diagnostics": PASSES_PER_STEP * npts * args.steps / (ms_nodiag * 1e-3),
                "phase_ms_per_step": {"advect_local_passes": phase[0], "rho+poisson": phase[1],
                                      "nccl_remap": phase[2], "diagnostics": phase[3],
                                      "advect_fused_remap_passes": phase[4] + phase[6],
                                      "barrier_after_fused_pass": phase[5] + phase[7],
                                      "fused_V_pass_x4+remap": phase[4], "barrier_after_V": phase[5],
                                      "fused_T_plane_x1+x2+rho+remap": phase[6],
                                      "allreduce_rho_after_T": phase[7],
                                      "how": f"separate {ph_steps}-step run with pooled CUDA events (sllb_set_phase_timers), not the timed region"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu_baseline, "check": check, "extra": {"c5_3d3v": c5}}
        print(json.dumps(line), flush=True)
    if comm is not None:
        comm.destroy()
        dist.destroy_process_group()


def main():
    # the image exports NCCL_DEBUG=VERSION, which makes NCCL print a banner on stdout; rank 0 must print ONE JSON line
    # NCCL prints its version banner on stdout at debug levels VERSION *and* WARN: drop the level the image exports and
    # send whatever else NCCL logs to stderr, so that stdout stays the one JSON line
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG", None)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank = env_int("RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    local_rank = env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
