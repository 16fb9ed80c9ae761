#!/bin/bash
# usage: profiles/scripts/r02_mgpu.sh N TAG   (under gpurun --gpus N): exactness checks, headline bench (with the C5 block) on N GPUs
N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29512 tests/mgpu/run_mgpu.py > gpurun_out/${TAG}_mgpu${N}.json 2> gpurun_out/${TAG}_mgpu${N}.err
timeout 500 $TR --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
for f in mgpu${N} bench_n${N}; do echo "== $f"; cat gpurun_out/${TAG}_$f.json | cut -c1-2500; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" gpurun_out/${TAG}_$f.err | tail -8; done
