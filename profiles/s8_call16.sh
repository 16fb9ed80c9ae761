mkdir -p gpurun_out
python -m pytest tests/test_gpu_stream.py tests/test_gpu_dd6d.py -q -x 2>&1 | tail -4
python profiles/bench_dd6d.py --steps 5 --warmup 2 > gpurun_out/s8_dd6d_n1b.json 2> gpurun_out/s8_dd6d_n1b.err
cut -c1-1500 gpurun_out/s8_dd6d_n1b.json; tail -3 gpurun_out/s8_dd6d_n1b.err
SLLB_PLANE=0 python profiles/bench_dd6d.py --steps 5 --warmup 2 > gpurun_out/s8_dd6d_n1c.json 2> gpurun_out/s8_dd6d_n1c.err
cut -c1-1500 gpurun_out/s8_dd6d_n1c.json
