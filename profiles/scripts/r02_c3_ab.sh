#!/bin/bash
# A/B of the single-GPU 2D2V step on the 64^4 (C3) and 128^4 (C4) grids: last reduction stage folded into the Poisson solve or not
export SLLB_SKIP_CPU=1 SLLB_SKIP_STREAM=1 SLLB_SKIP_C5=1
for fold in 0 1; do
  for n in 64 128; do
    SLLB_FOLD_SUMS=$fold SLLB_BENCH_N=$n python bench.py --steps 50 --warmup 5 > gpurun_out/r02s8_n${n}_fold${fold}.json 2> gpurun_out/r02s8_n${n}_fold${fold}.err
    python - <<PY
import json
d=json.load(open('gpurun_out/r02s8_n${n}_fold${fold}.json'))
print('n=${n} fold=${fold}', round(d['ms_per_step'],4), round(d['ms_per_step_without_diagnostics'],4), d['gpu_launches']/d['steps'], {k:round(v,4) for k,v in d['phase_ms_per_step'].items() if isinstance(v,float) and v>0})
PY
  done
done
SLLB_FOLD_SUMS=1 SLLB_BENCH_N=64 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02s8_launches_c3_fold1.csv python bench.py --steps 4 --warmup 3 > /dev/null 2>&1
SLLB_FOLD_SUMS=0 SLLB_BENCH_N=64 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02s8_launches_c3_fold0.csv python bench.py --steps 4 --warmup 3 > /dev/null 2>&1
python -m pytest tests/test_gpu_sim2d_nml.py -q 2>&1 | tail -3
