mkdir -p gpurun_out
python -m pytest tests/test_gpu_stream.py -q 2>&1 | tail -20 > gpurun_out/s8_stream_tests.log
tail -12 gpurun_out/s8_stream_tests.log
python bench.py --steps 30 --warmup 5 > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err
tail -3 gpurun_out/s8_bench.err; cut -c1-6000 gpurun_out/s8_bench.json
