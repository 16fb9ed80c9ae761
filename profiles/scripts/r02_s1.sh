#!/bin/bash
# round 2, session 1: evidence at the round-1 HEAD: ncu --set full of the T-stage plane kernel and the two V-stage kernels,
# DRAM traffic of the bench command per kernel, compute-sanitizer memcheck + racecheck on smoke() and a small 2D2V run.
O=gpurun_out
export SLLB_SKIP_CPU=1 SLLB_SKIP_STREAM=1 SLLB_E2E_STEPS=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/r02s1_smi.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spline_plane_r|k_spline_strided_split|k_spline_contig_split' -c 6 -o $O/r02s1_full -f python profiles/prof_kernels.py > $O/r02s1_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/r02s1_launches_traffic.csv python bench.py --steps 2 --warmup 1 > $O/r02s1_b.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file $O/r02s1_memcheck_smoke.log python __graft_entry__.py --smoke > $O/r02s1_memcheck_smoke.out 2>&1
timeout 900 compute-sanitizer --tool racecheck --log-file $O/r02s1_racecheck_smoke.log python __graft_entry__.py --smoke > $O/r02s1_racecheck_smoke.out 2>&1
timeout 900 compute-sanitizer --tool racecheck --log-file $O/r02s1_racecheck_sim4d.log python profiles/scripts/sanitize_sim4d.py > $O/r02s1_racecheck_sim4d.out 2>&1
timeout 600 compute-sanitizer --tool memcheck --log-file $O/r02s1_memcheck_sim4d.log python profiles/scripts/sanitize_sim4d.py > $O/r02s1_memcheck_sim4d.out 2>&1
tail -3 $O/r02s1_*smoke.out $O/r02s1_*sim4d.out; tail -5 $O/r02s1_*check*.log
python bench.py --steps 20 --warmup 5 > $O/r02s1_bench.json 2> $O/r02s1_bench.err; cut -c1-1500 $O/r02s1_bench.json
