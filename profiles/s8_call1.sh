mkdir -p gpurun_out
python -m pytest tests/test_gpu_spline_dd.py -x -q 2>&1 | tail -40 > gpurun_out/s8_spline_tests.log
python profiles/prof_spline_dd.py > gpurun_out/s8_spline_dd_perf.json 2> gpurun_out/s8_spline_dd_perf.err
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/s8_gpu_tests.log
tail -5 gpurun_out/s8_spline_tests.log; cat gpurun_out/s8_spline_dd_perf.json | cut -c1-3000; tail -3 gpurun_out/s8_gpu_tests.log
