mkdir -p gpurun_out
SLLB_C2_N=128 SLLB_C2_METHOD=spline python profiles/bench_c2.py > gpurun_out/s8_c1.json 2> gpurun_out/s8_c1.err
cat gpurun_out/s8_c1.json
SLLB_BENCH_N=64 SLLB_SKIP_CPU=1 python bench.py --steps 100 --warmup 5 > gpurun_out/s8_c3.json 2> gpurun_out/s8_c3.err
cut -c1-330 gpurun_out/s8_c3.json; tail -2 gpurun_out/s8_c3.err
