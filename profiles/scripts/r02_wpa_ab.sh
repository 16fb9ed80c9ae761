#!/bin/bash
# T-stage plane kernel: bulk-copy pass A (SLLB_PLANE_WPA=0) vs warp-cooperative pass A from global memory (=1):
# parity tests under both, then CUDA-event timings of the kernel alone and the 128^4 step
mkdir -p gpurun_out
SLLB_PLANE_WPA=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_sim4d.py -m gpu -x -q -k "plane or sim4d or remap" 2>&1 | tail -5
for wp in 0 1; do for pf in 0 2; do echo "== SLLB_PLANE_WPA=$wp SLLB_PLANE_L2_PREFETCH=$pf"; SLLB_PLANE_WPA=$wp SLLB_PLANE_L2_PREFETCH=$pf timeout 300 python profiles/ab_plane.py 2>&1 | grep "ept= 0\|rror"; done; done
for wp in 0 1; do echo "== bench SLLB_PLANE_WPA=$wp"; SLLB_PLANE_WPA=$wp timeout 600 SLLB_SKIP_CPU=1 SLLB_SKIP_C5=1 SLLB_SKIP_STREAM=1 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d.get('extra',{}).get('phases_ms'))"; done
