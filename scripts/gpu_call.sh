mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 1 --no-secondary > gpurun_out/r2g_nccl_n2.json 2> gpurun_out/r2g_nccl_n2.log; echo rc=$?
grep -E "NCCL INFO (Channel|Connected|comm|NVLS|Using|Trees|Ring|AllGather|AllReduce|Broadcast)" gpurun_out/r2g_nccl_n2.log | grep -v "Channel [0-9][0-9]/" | head -60 > gpurun_out/r2g_nccl_n2_summary.log; wc -l gpurun_out/r2g_nccl_n2_summary.log; tail -5 gpurun_out/r2g_nccl_n2_summary.log
