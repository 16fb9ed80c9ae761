mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29513 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/s8_bench_n8.json 2> gpurun_out/s8_bench_n8.err
echo rc=$?
cut -c1-3500 gpurun_out/s8_bench_n8.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" gpurun_out/s8_bench_n8.err | tail -5
