#!/bin/bash
# T-stage plane kernel with / without the L2 prefetch of the next plane (CUDA events, 20 launches each, 128^4)
for pf in 0 1; do echo "== SLLB_PLANE_L2_PREFETCH=$pf"; SLLB_PLANE_L2_PREFETCH=$pf python profiles/ab_plane.py 2>&1 | grep "ept= 0\|separate"; done
