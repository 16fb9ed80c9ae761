mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/s8_gpu_tests7.log
tail -4 gpurun_out/s8_gpu_tests7.log
python profiles/bench_dd6d.py --steps 5 --warmup 2 > gpurun_out/s8_dd6d_n1.json 2> gpurun_out/s8_dd6d_n1.err
cut -c1-1500 gpurun_out/s8_dd6d_n1.json; tail -3 gpurun_out/s8_dd6d_n1.err
