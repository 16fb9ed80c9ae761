"""world_size-2 gloo test of the slim 6D halo exchange host logic (neighbours, which planes go where):
mirrors src/parallelization/decomposition/testing/test_decomposition_slim.F90:113-130 -- after the exchange
along each axis with halo widths 1..3 the halo buffers hold the periodic neighbours' planes.  The plan comes
from the C ABI (sllb_dd6d_plan); the exchange pattern is the one sllb_dd6d_halo_exchange runs with NCCL
(first hw_right planes -> left neighbour, last hw_left planes -> right neighbour)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, g, procs, result):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import selalib_b200 as sb
    lay = sb.dd6d_plan(world, rank, g, procs)
    glob = np.arange(np.prod(g), dtype=np.float64).reshape(g, order="F")
    sl = tuple(slice(lay["mn"][d], lay["mn"][d] + lay["nw"][d]) for d in range(6))
    local = glob[sl].copy(order="F")
    ok = True
    for axis in range(6):
        n = lay["nw"][axis]
        for hwl, hwr in ((1, 1), (2, 3), (3, 1)):
            if max(hwl, hwr) > n:
                continue
            take = lambda a, lo, hi: np.ascontiguousarray(np.take(a, range(lo, hi), axis=axis).ravel(order="F"))
            send_lo, send_hi = take(local, 0, hwr), take(local, n - hwl, n)
            shp_l = list(lay["nw"]); shp_l[axis] = hwl
            shp_r = list(lay["nw"]); shp_r[axis] = hwr
            if lay["procs"][axis] == 1:
                halo_r, halo_l = send_lo, send_hi
            else:
                halo_r = torch.empty(send_lo.size, dtype=torch.float64)
                halo_l = torch.empty(send_hi.size, dtype=torch.float64)
                reqs = [dist.isend(torch.from_numpy(send_lo), lay["left"][axis]),
                        dist.irecv(halo_r, lay["right"][axis]),
                        dist.isend(torch.from_numpy(send_hi), lay["right"][axis]),
                        dist.irecv(halo_l, lay["left"][axis])]
                for q in reqs:
                    q.wait()
                halo_r, halo_l = halo_r.numpy(), halo_l.numpy()
            # expected: periodic continuation of the global array beyond my block
            idx_r = [(lay["mn"][axis] + n + j) % g[axis] for j in range(hwr)]
            idx_l = [(lay["mn"][axis] - hwl + j) % g[axis] for j in range(hwl)]
            others = tuple(s for d, s in enumerate(sl) if d != axis)

            def expect(idx):
                e = np.take(glob, idx, axis=axis)
                sel = list(sl); sel[axis] = slice(None)
                return e[tuple(sel)]
            ok = ok and np.array_equal(halo_r.reshape(shp_r, order="F"), expect(idx_r))
            ok = ok and np.array_equal(halo_l.reshape(shp_l, order="F"), expect(idx_l))
    result[rank] = 1 if ok else 0
    dist.destroy_process_group()


@pytest.mark.parametrize("g,procs", [([4, 4, 4, 4, 4, 6], None), ([4, 4, 4, 4, 6, 4], [1, 1, 1, 1, 2, 1]),
                                     ([4, 6, 4, 4, 4, 4], [1, 2, 1, 1, 1, 1])])
def test_halo_exchange_world2(g, procs):
    world = 2
    result = mp.Array("i", [0] * world)
    port = _free_port()
    ps = [mp.Process(target=_worker, args=(r, world, port, g, procs, result)) for r in range(world)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    assert list(result) == [1] * world


def test_plan_matches_reference_process_grid():
    import selalib_b200 as sb
    # sll_f_set_process_grid table (sll_m_decomposition.F90:2489-2498)
    assert sb.dd6d_plan(8, 0, [8] * 6)["procs"] == (1, 1, 1, 2, 2, 2)
    assert sb.dd6d_plan(2, 1, [8] * 6)["procs"] == (1, 1, 1, 1, 1, 2)
    lay = sb.dd6d_plan(8, 5, [8] * 6)   # MPI_Cart_create order: last axis fastest -> coords (.,.,.,1,0,1)
    assert lay["coords"] == (0, 0, 0, 1, 0, 1) and lay["mn"] == (0, 0, 0, 4, 0, 4)
    assert lay["left"][5] == 4 and lay["right"][5] == 4 and lay["left"][3] == 1
    with pytest.raises(sb.SllbError):
        sb.dd6d_plan(2, 0, [8, 8, 8, 8, 8, 7])
