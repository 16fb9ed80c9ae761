mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for ch in 1 4; do
SLLB_HALO_CHUNKS=$ch timeout 300 $TR --master-port 2951$ch profiles/bench_dd6d.py --steps 5 --warmup 2 > gpurun_out/s8_dd6d_n2_c$ch.json 2> gpurun_out/s8_dd6d_n2_c$ch.err
python - <<PY
import json
l=[x for x in open('gpurun_out/s8_dd6d_n2_c$ch.json').read().splitlines() if x.startswith('{')]
r=json.loads(l[0]) if l else {}
print('chunks $ch', {k:r.get(k) for k in ('value','ms_per_step','x_pass_ms','v_pass_ms','halo_ms_per_split_pass','mass')})
PY
done
