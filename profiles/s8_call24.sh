mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29514 profiles/bench_dd6d.py --steps 5 --warmup 2 > gpurun_out/s8_dd6d_n8.json 2> gpurun_out/s8_dd6d_n8.err
cut -c1-1500 gpurun_out/s8_dd6d_n8.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/s8_dd6d_n8.err | tail -4
SLLB_HALO_CHUNKS=1 timeout 300 $TR --master-port 29515 profiles/bench_dd6d.py --steps 5 --warmup 2 > gpurun_out/s8_dd6d_n8_c1.json 2> gpurun_out/s8_dd6d_n8_c1.err
cut -c1-1500 gpurun_out/s8_dd6d_n8_c1.json
