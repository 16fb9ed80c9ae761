mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/s8_gpu_tests4.log
tail -3 gpurun_out/s8_gpu_tests4.log
python profiles/prof_spline_dd.py > gpurun_out/s8_spline_dd_perf2.json 2> gpurun_out/s8_spline_dd_perf2.err
python - <<'PY'
import json
r=json.load(open('gpurun_out/s8_spline_dd_perf2.json'))
for k,v in r['kernels'].items():
    print(f"{k:55s}", {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
