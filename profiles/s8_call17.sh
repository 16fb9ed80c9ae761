mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'lagrange_plane' -c 1 -o gpurun_out/s8_lplane python profiles/prof_lagrange_plane.py > gpurun_out/s8_ncu_lplane.log 2>&1
tail -2 gpurun_out/s8_ncu_lplane.log
