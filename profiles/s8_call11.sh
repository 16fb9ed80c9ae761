mkdir -p gpurun_out
python profiles/bench_c2.py > gpurun_out/s8_c2.json 2> gpurun_out/s8_c2.err
cat gpurun_out/s8_c2.json; tail -3 gpurun_out/s8_c2.err
