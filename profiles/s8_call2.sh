mkdir -p gpurun_out
python -m pytest tests/test_gpu_spline_dd.py -q 2>&1 | tail -40 > gpurun_out/s8_spline_tests2.log
tail -8 gpurun_out/s8_spline_tests2.log
