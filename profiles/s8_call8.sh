mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'contig|spline_dd' -c 5 -o gpurun_out/s8_contig python profiles/prof_contig.py > gpurun_out/s8_ncu_contig.log 2>&1
tail -3 gpurun_out/s8_ncu_contig.log
ls -la gpurun_out/*.ncu-rep
