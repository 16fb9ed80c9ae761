#!/bin/bash
# plane kernel, default policy (compile-time extents + TMEM accumulators where they apply): variant tests, kernel timings,
# steps at 128^4 and 64^4 with the accumulators in TMEM / registers + shared memory, ncu --set full of the fused-density launch
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plane" 2>&1 | tail -2
echo "== defaults"; SLLB_AB_EPT0_ONLY=1 timeout 300 python profiles/ab_plane.py 2>&1 | grep "ept= 0\|rror"
for n in 128 64; do for tm in 0 1; do echo "== bench N=$n SLLB_PLANE_TMEM=$tm"; SLLB_BENCH_N=$n SLLB_PLANE_TMEM=$tm SLLB_SKIP_CPU=1 SLLB_SKIP_C5=1 SLLB_SKIP_STREAM=1 timeout 600 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['roofline']['t_stage_plane_kernel']['ms_per_launch'], d['check']['mass'], d['check']['field_energy'], d['check']['checksum_wf2'])"; done; done
SLLB_AB_EPT0_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spline_plane_r<\(bool\)1' -s 4 -c 1 -o gpurun_out/r02_plane_tmem_full -f python profiles/ab_plane.py > gpurun_out/r02_plane_tmem_full.log 2>&1; tail -2 gpurun_out/r02_plane_tmem_full.log
