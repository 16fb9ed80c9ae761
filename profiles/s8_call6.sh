mkdir -p gpurun_out
python -m pytest tests/test_gpu_splitting.py -q -k "namelist or STRANG or LIE" 2>&1 | tail -40 > gpurun_out/s8_nml_tests.log
tail -30 gpurun_out/s8_nml_tests.log
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/s8_gpu_tests3.log
tail -5 gpurun_out/s8_gpu_tests3.log
