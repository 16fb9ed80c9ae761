"""world_size-2 gloo test of the multi-rank remap path (host logic): each rank packs the boxes the plan
(sllb_remap4d_plan) assigns to every peer in column-major order, exchanges them with grouped send/recv and
unpacks -- the same pack / exchange / unpack sequence the CUDA path runs with NCCL
(sllb_dist4d_remap; reference apply_remap_4D_double, sll_m_remapper.F90:3308-3456)."""
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


def _worker(rank, world, port, g, result):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import selalib_b200 as sb
    f1, f2 = sb.factorize_in_two_powers_of_two(world)
    layouts = {0: [1, 1, f1, f2], 1: [f1, f2, 1, 1]}
    glob = np.arange(np.prod(g), dtype=np.float64).reshape(g, order="F")
    ok = True
    local = None
    for direction in (0, 1):                 # x-seq -> v-seq, then back
        pf, pt = layouts[direction], layouts[1 - direction]
        bf = sb.layout4d_boxes(g, pf, world)[rank]
        bt = sb.layout4d_boxes(g, pt, world)[rank]
        if local is None:
            local = glob[tuple(slice(bf[d, 0], bf[d, 1] + 1) for d in range(4))].copy(order="F")
        sboxes, rboxes = sb.remap4d_plan(g, pf, pt, world, rank)
        send, recv = [], []
        for r in range(world):
            s = sboxes[r]
            blk = local[tuple(slice(s[d, 0] - bf[d, 0], s[d, 1] + 1 - bf[d, 0]) for d in range(4))]
            send.append(torch.from_numpy(np.ascontiguousarray(blk.ravel(order="F"))))
            rb = rboxes[r]
            recv.append(torch.empty(int(np.prod([max(0, rb[d, 1] - rb[d, 0] + 1) for d in range(4)])), dtype=torch.float64))
        # grouped send/recv, the same pattern as the ncclSend/ncclRecv group of the CUDA path
        recv[rank].copy_(send[rank])
        reqs = []
        for r in range(world):
            if r != rank:
                if send[r].numel():
                    reqs.append(dist.isend(send[r], r))
                if recv[r].numel():
                    reqs.append(dist.irecv(recv[r], r))
        for q in reqs:
            q.wait()
        new = np.full([bt[d, 1] - bt[d, 0] + 1 for d in range(4)], -1.0, order="F")
        for r in range(world):
            rb = rboxes[r]
            shp = [rb[d, 1] - rb[d, 0] + 1 for d in range(4)]
            new[tuple(slice(rb[d, 0] - bt[d, 0], rb[d, 1] + 1 - bt[d, 0]) for d in range(4))] = \
                recv[r].numpy().reshape(shp, order="F")
        ok = ok and np.array_equal(new, glob[tuple(slice(bt[d, 0], bt[d, 1] + 1) for d in range(4))])
        local = new
    # rho tiles gathered to every rank (split_to_full): allgather + placement by the v-layout boxes
    bv = sb.layout4d_boxes(g, layouts[1], world)
    tile = glob[:, :, 0, 0][tuple(slice(bv[rank, d, 0], bv[rank, d, 1] + 1) for d in range(2))]
    tiles = [torch.empty(tile.size, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(tiles, torch.from_numpy(np.ascontiguousarray(tile.ravel(order="F"))))
    full = np.zeros(g[:2])
    for r in range(world):
        shp = [bv[r, d, 1] - bv[r, d, 0] + 1 for d in range(2)]
        full[tuple(slice(bv[r, d, 0], bv[r, d, 1] + 1) for d in range(2))] = tiles[r].numpy().reshape(shp, order="F")
    ok = ok and np.array_equal(full, glob[:, :, 0, 0])
    result[rank] = 1 if ok else 0
    dist.destroy_process_group()


@pytest.mark.parametrize("g", [[8, 8, 8, 8], [9, 8, 10, 7]])
def test_remap_roundtrip_world2(g):
    world = 2
    result = mp.Array("i", [0] * world)
    port = _free_port()
    procs = [mp.Process(target=_worker, args=(r, world, port, g, result)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(result) == [1] * world
