#!/bin/bash
# fused-diagnostics variant of the strided spline pass after moving its kinetic weights to shared memory:
# parity (simulation tests), then the 128^4 step
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_baseline_sizes.py tests/test_gpu_parity.py -m gpu -x -q -k "sim4d or diag or thdiag or 2d2v or landau" 2>&1 | tail -3
SLLB_SKIP_CPU=1 SLLB_SKIP_C5=1 SLLB_SKIP_STREAM=1 timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_diag_ab_bench.json 2> gpurun_out/r02_diag_ab_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_diag_ab_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d.get('value_without_diagnostics'), d['roofline']['traffic'], d['roofline']['frac'], d['roofline']['t_stage_plane_kernel']['ms_per_launch'])
print(d.get('extra',{}).get('phase_ms_per_step'))
print(d['check'])
PY
