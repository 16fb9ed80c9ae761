mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/s8_gpu_tests6.log
tail -4 gpurun_out/s8_gpu_tests6.log
python profiles/bench_c2.py > gpurun_out/s8_c2c.json 2> gpurun_out/s8_c2c.err
cat gpurun_out/s8_c2c.json; tail -3 gpurun_out/s8_c2c.err
