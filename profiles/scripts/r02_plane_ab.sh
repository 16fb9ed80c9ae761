#!/bin/bash
# T-stage plane kernel (CUDA events, 20 launches each, 128^4): charge-density accumulators in registers + shared memory vs in
# tensor memory, with / without the L2 prefetch of the next plane
for tm in 0 1; do for pf in 0 2; do echo "== SLLB_PLANE_TMEM=$tm SLLB_PLANE_L2_PREFETCH=$pf"; SLLB_PLANE_TMEM=$tm SLLB_PLANE_L2_PREFETCH=$pf timeout 300 python profiles/ab_plane.py 2>&1 | grep "ept= 0\|separate\|rror"; done; done
