#!/bin/bash
# round 2, final single-GPU evidence at HEAD: GPU suite, bench, launch list + DRAM traffic of the same bench command,
# ncu --set full of the hot kernels, compute-sanitizer memcheck + racecheck on smoke() and small simulations
O=gpurun_out; T=${1:-r02f}
python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -4 > $O/${T}_gpu_tests.log; cat $O/${T}_gpu_tests.log
python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; cut -c1-260 $O/${T}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; cut -c1-200 $O/${T}_bench_ref.json
SLLB_SKIP_CPU=1 SLLB_SKIP_STREAM=1 SLLB_SKIP_C5=1 SLLB_E2E_STEPS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 500 --csv --log-file $O/${T}_launches_traffic.csv python bench.py --steps 3 --warmup 3 > $O/${T}_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_spline_plane_r|k_spline_strided_split|k_spline_contig_split|k_pd_|k_reduce_stage' -c 12 -o $O/${T}_full -f python profiles/prof_kernels.py > $O/${T}_full.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file $O/${T}_memcheck_smoke.log python __graft_entry__.py --smoke > $O/${T}_memcheck_smoke.out 2>&1
timeout 900 compute-sanitizer --tool racecheck --log-file $O/${T}_racecheck_smoke.log python __graft_entry__.py --smoke > $O/${T}_racecheck_smoke.out 2>&1
timeout 900 compute-sanitizer --tool racecheck --log-file $O/${T}_racecheck_sims.log python profiles/scripts/sanitize_sim4d.py > $O/${T}_racecheck_sims.out 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file $O/${T}_memcheck_sims.log python profiles/scripts/sanitize_sim4d.py > $O/${T}_memcheck_sims.out 2>&1
for f in $O/${T}_*check*.log; do echo "$f: $(tail -n 1 $f)"; done; for f in $O/${T}_*check*.out; do echo "$f: $(tail -n 1 $f)"; done
