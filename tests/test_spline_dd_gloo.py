"""world_size-2 gloo test of the local-spline split-axis protocol (sllb_dd6d_advect_axis_spline): which rank computes
which part of the two boundary sums, and where the sums and the halo cells travel.

Reference: sll_s_advection_6d_spline_dd_slim_advect_eta4 (src/semi_lagrangian/advection/
sll_m_advection_6d_spline_dd_slim.F90:1030-1100): prepare_exchange on my piece -> bc exchange + halo exchange of one cell
per side -> finish_boundary_conditions + interpolant + eval on my piece.  Every rank runs the per-line device functions of
the CUDA path (sllb_spline15.cuh compiled for the host, tests/host/) on ITS piece of a batch of lines and exchanges with its
ring neighbours over gloo exactly as the library does over NVLink / NCCL:
    for_right (my top cells' part of d_0)      -> right neighbour's bc_left
    for_left  (my bottom cells' part of c_np2) -> left  neighbour's bc_right
    my first hw_right cells -> left neighbour's right halo,  my last hw_left cells -> right neighbour's left halo.
The gathered result must equal the oracle's emulation of the same decomposition to 1e-12."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, nlines, npiece, hw, datafile, result):
    # NOTE: the forked children must not enter the oracle (OpenMP after fork deadlocks when the parent already ran a
    # parallel region): the parent computes the reference and hands it over in `datafile`
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
    from host import emu
    data = np.load(datafile)
    glob, disp, ref_all = data["glob"], data["disp"], data["ref"]   # every rank sees the same global lines
    n = world * npiece
    left, right = (rank - 1) % world, (rank + 1) % world
    mine = glob[:, rank * npiece:(rank + 1) * npiece].copy()
    si = np.floor(disp).astype(int)
    # K9p on my piece
    sums = np.array([emu.spline_dd_prepare(mine[l], int(si[l])) for l in range(nlines)])
    for_right, for_left = np.ascontiguousarray(sums[:, 0]), np.ascontiguousarray(sums[:, 1])
    bc_left, bc_right = torch.empty(nlines, dtype=torch.float64), torch.empty(nlines, dtype=torch.float64)
    halo_l = torch.empty(nlines * hw[0], dtype=torch.float64)
    halo_r = torch.empty(nlines * hw[1], dtype=torch.float64)
    send_lo = np.ascontiguousarray(mine[:, :hw[1]]).ravel()
    send_hi = np.ascontiguousarray(mine[:, npiece - hw[0]:]).ravel()
    reqs = [dist.isend(torch.from_numpy(for_right), right, tag=1), dist.irecv(bc_left, left, tag=1),
            dist.isend(torch.from_numpy(for_left), left, tag=2), dist.irecv(bc_right, right, tag=2),
            dist.isend(torch.from_numpy(send_lo), left, tag=3), dist.irecv(halo_r, right, tag=3),
            dist.isend(torch.from_numpy(send_hi), right, tag=4), dist.irecv(halo_l, left, tag=4)]
    for q in reqs:
        q.wait()
    hl = halo_l.numpy().reshape(nlines, hw[0]); hr = halo_r.numpy().reshape(nlines, hw[1])
    out = np.stack([emu.spline_dd_piece(mine[l], hl[l], hr[l], int(si[l]), float(disp[l] - si[l]), float(bc_left[l]),
                                        float(bc_right[l])) for l in range(nlines)])
    # the oracle's emulation of this decomposition, my piece of it
    ref = ref_all[:, rank * npiece:(rank + 1) * npiece]
    err = np.abs(out - ref).max() / np.abs(glob).max()
    result[rank] = 1 if err <= 1e-12 else 0
    dist.destroy_process_group()


def _run(tmp_path, world, nlines, npiece, hw):
    from host import emu
    from oracle import orc
    emu.lib()                                                # build the host library once, in the parent
    rng = np.random.default_rng(20261017)
    n = world * npiece
    glob = rng.standard_normal((nlines, n))
    disp = rng.uniform(-hw[0], hw[1], nlines) * 0.999       # shifts in [-hw_left, hw_right - 1]
    f = np.asfortranarray(glob.reshape(nlines, n, 1).copy())
    ref = orc.spline_dd_advect_axis(f, 1, world, disp, (1, 1, 0, 1, nlines, 1))[:, :, 0]
    datafile = str(tmp_path / "lines.npz")
    np.savez(datafile, glob=glob, disp=disp, ref=ref)
    result = mp.Array("i", [0] * world)
    port = _free_port()
    ps = [mp.Process(target=_worker, args=(r, world, port, nlines, npiece, hw, datafile, result)) for r in range(world)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(120)
        if p.is_alive():
            p.kill()
        assert p.exitcode == 0
    assert list(result) == [1] * world


def test_spline_dd_exchange_world2(tmp_path):
    _run(tmp_path, 2, 12, 20, (1, 1))       # the eta4..6 case: one halo cell per side, shifts 0 and -1


def test_spline_dd_exchange_world3_wider_halos(tmp_path):
    _run(tmp_path, 3, 9, 24, (2, 3))        # three ranks tell left from right; shifts -2 .. 2
