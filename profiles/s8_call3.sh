mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/mgpu/run_mgpu.py > gpurun_out/s8_mgpu2.json 2> gpurun_out/s8_mgpu2.err
echo rc=$?
cat gpurun_out/s8_mgpu2.json | cut -c1-3000; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" gpurun_out/s8_mgpu2.err | tail -15
