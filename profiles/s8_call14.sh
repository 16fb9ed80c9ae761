mkdir -p gpurun_out
python bench.py --steps 30 --warmup 5 > gpurun_out/s8_bench_final.json 2> gpurun_out/s8_bench_final.err
cut -c1-400 gpurun_out/s8_bench_final.json
SLLB_SKIP_CPU=1 SLLB_SKIP_STREAM=1 SLLB_E2E_STEPS=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s8_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/s8_b.log 2>&1
tail -2 gpurun_out/s8_b.log | cut -c1-200
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s8_bench_ref.json 2> gpurun_out/s8_bench_ref.err
cut -c1-600 gpurun_out/s8_bench_ref.json
