#!/bin/bash
# T-stage plane kernel: generic extents (SLLB_PLANE_CONST_DIMS=0) vs the 128 x 128 / 64 x 64 instantiations with compile-time
# extents (=1): parity tests, kernel alone (CUDA events, 20 launches, 128^4), then the 128^4 and 64^4 steps
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py tests/test_gpu_baseline_sizes.py -m gpu -x -q -k "plane or sim4d or remap or 2d2v or landau" 2>&1 | tail -3
for cd in 0 1; do for pf in 0 1 2; do echo "== SLLB_PLANE_CONST_DIMS=$cd SLLB_PLANE_L2_PREFETCH=$pf"; SLLB_AB_EPT0_ONLY=1 SLLB_PLANE_CONST_DIMS=$cd SLLB_PLANE_L2_PREFETCH=$pf timeout 300 python profiles/ab_plane.py 2>&1 | grep "ept= 0\|rror"; done; done
for n in 128 64; do for cd in 0 1; do echo "== bench N=$n SLLB_PLANE_CONST_DIMS=$cd"; SLLB_BENCH_N=$n SLLB_PLANE_CONST_DIMS=$cd SLLB_SKIP_CPU=1 SLLB_SKIP_C5=1 SLLB_SKIP_STREAM=1 timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['t_stage_plane_kernel']['ms_per_launch'])"; done; done
