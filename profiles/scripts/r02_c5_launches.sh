#!/bin/bash
# 3D3V 32^6 on one GPU: launch list (time + DRAM bytes per kernel) of three time steps
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_c5_launches.csv python profiles/bench_dd6d.py --steps 2 --warmup 1 > gpurun_out/r02_c5_launches.log 2>&1
python profiles/scripts/summarize_ncu.py list gpurun_out/r02_c5_launches.csv gpurun_out/r02_c5_launches.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_c5_launches.json'))
for k,v in d['kernels'].items():
    print(f"{k[:70]:70s} n={v['launches']:4d} us={v['us_per_launch']:9.1f} share={v['share_of_kernel_time']:.3f} dram={v['dram_bytes_per_launch']/1e9:.3f}")
PY
