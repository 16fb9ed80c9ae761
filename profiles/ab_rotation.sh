#!/bin/bash
# usage: profiles/ab_rotation.sh N TAG  (under gpurun --gpus N): fused-remap A/B, rank-rotated sweep on/off
N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for rot in 1 0; do
  SLLB_E2E_STEPS=2 SLLB_REMAP_ROTATION=$rot timeout 200 $TR --master-port 2951$rot bench.py --gpus $N --steps 30 --warmup 5 \
     > gpurun_out/${TAG}_bench_n${N}_rot${rot}.json 2> gpurun_out/${TAG}_bench_n${N}_rot${rot}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_n${N}_rot${rot}.json").read().strip().splitlines()[-1])
    print("rot=${rot}", "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], json.dumps(d["phase_ms_per_step"]))
except Exception as e:
    print("rot=${rot} failed", e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" gpurun_out/${TAG}_bench_n${N}_rot${rot}.err | tail -5
done
