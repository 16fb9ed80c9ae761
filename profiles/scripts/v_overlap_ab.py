"""A/B of the chunked V stage of the multi-GPU 2D2V loop (torchrun, N ranks): ms per step of sllb_sim4d_run on 128^4 with
the x3 pass overlapped under the x4 + remap pass (2, 4, 8 chunks) and without."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import selalib_b200 as sb  # noqa: E402

if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "WARN"):
    os.environ.pop("NCCL_DEBUG", None)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
sb.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.tensor(list(sb.Comm.unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
comm = sb.Comm(bytes(idt.cpu().tolist()), world, rank)
n = int(os.environ.get("SLLB_BENCH_N", "128"))
res = {}
for tag, ov in (("whole", 0), ("chunks2", 2), ("chunks4", 4), ("chunks8", 8), ("whole_again", 0)):
    sb.set_v_overlap(ov)
    S = sb.Sim4d([n] * 4, [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1, comm=comm)
    S.run(5, diagnostics=False)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    S.run(30, diagnostics=False)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 30], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[tag] = float(t.item())
    S.destroy()
if rank == 0:
    print(json.dumps({"n_gpus": world, "ms_per_step": res}), flush=True)
comm.destroy()
dist.destroy_process_group()
