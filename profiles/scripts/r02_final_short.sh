#!/bin/bash
# after the last kernel change of the round: GPU suite, bench line, reference arm, 3D3V launch list
O=gpurun_out; T=${1:-r02f5}
python -m pytest tests -m gpu -q 2>&1 | grep -v "^$" | tail -4 > $O/${T}_gpu_tests.log; cat $O/${T}_gpu_tests.log
python bench.py --steps 20 --warmup 5 > $O/${T}_bench.json 2> $O/${T}_bench.err; cut -c1-260 $O/${T}_bench.json
python __graft_entry__.py --smoke 2>&1 | tail -1
bash profiles/scripts/r02_c5_launches.sh 2>&1 | tail -16 > $O/${T}_c5_launches.txt; cat $O/${T}_c5_launches.txt
