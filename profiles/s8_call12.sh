mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/s8_gpu_tests5.log
tail -4 gpurun_out/s8_gpu_tests5.log
python profiles/bench_c2.py > gpurun_out/s8_c2b.json 2> gpurun_out/s8_c2b.err
cat gpurun_out/s8_c2b.json; tail -3 gpurun_out/s8_c2b.err
