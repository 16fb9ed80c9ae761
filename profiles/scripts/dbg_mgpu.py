import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "/root/repo")
import selalib_b200 as sb
os.environ.pop("NCCL_DEBUG", None)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); sb.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0: idt.copy_(torch.tensor(list(sb.Comm.unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
comm = sb.Comm(bytes(idt.cpu().tolist()), world, rank)
a4 = ([32, 32, 32, 32], [0, 0, -6, -6], [4 * np.pi, 4 * np.pi, 6, 6], 0.5, 0.5, 1e-3, 0.1)
out = {}
for tag, kw in (("1gpu_graph", dict(g=True)), ("1gpu_nograph", dict(g=False))):
    sb.set_cuda_graphs(kw["g"])
    S = sb.Sim4d(*a4); r = S.run(3); f = S.field().download(); S.destroy()
    out[tag] = (r, f)
sb.set_cuda_graphs(True)
for tag, env in (("2gpu", {}), ("2gpu_nodirect", {"pd": False})):
    if "pd" in env: sb.set_poisson_direct(False)
    S = sb.Sim4d(*a4, comm=comm); r = S.run(3); f = S.field().download(); b = S.box(0); S.destroy()
    sb.set_poisson_direct(True)
    out[tag] = (r, f, b)
if rank == 0:
    r1, f1 = out["1gpu_nograph"]
    print("graph vs nograph rows", np.abs(out["1gpu_graph"][0] / r1 - 1).max(axis=0), "f", np.abs(out["1gpu_graph"][1] - f1).max())
    for tag in ("2gpu", "2gpu_nodirect"):
        r, f, b = out[tag]
        sx = tuple(slice(b[d, 0], b[d, 1] + 1) for d in range(4))
        print(tag, "rows rel per col", np.abs(r / r1 - 1).max(axis=0), "f", np.abs(f - f1[sx]).max() / np.abs(f1).max())
        print(r); print(r1)
comm.destroy(); dist.destroy_process_group()
