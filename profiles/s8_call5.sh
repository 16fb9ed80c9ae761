mkdir -p gpurun_out
python -m pytest tests/test_gpu_hermite.py -q 2>&1 | tail -40 > gpurun_out/s8_hermite_tests.log
tail -30 gpurun_out/s8_hermite_tests.log
