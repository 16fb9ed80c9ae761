#!/bin/bash
# plane kernel with compile-time extents: accumulators in registers + shared memory vs tensor memory; then one ncu --set full
# capture of the fused-density instantiation
for tm in 0 1; do echo "== SLLB_PLANE_TMEM=$tm (const dims, default prefetch policy)"; SLLB_AB_EPT0_ONLY=1 SLLB_PLANE_TMEM=$tm timeout 300 python profiles/ab_plane.py 2>&1 | grep "ept= 0\|rror"; done
SLLB_PLANE_TMEM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plane" 2>&1 | tail -2
SLLB_AB_EPT0_ONLY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spline_plane_r' -s 6 -c 2 -o gpurun_out/r02_plane_const_full -f python profiles/ab_plane.py > gpurun_out/r02_plane_const_full.log 2>&1; tail -2 gpurun_out/r02_plane_const_full.log
