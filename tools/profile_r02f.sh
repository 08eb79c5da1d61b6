#!/bin/bash
# peer-memory exchange kernel vs NCCL path, 2 GPUs
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/check_sharded.py > gpurun_out/r02f_check_sharded_n2.log 2>&1
echo "rc=$?"; tail -15 gpurun_out/r02f_check_sharded_n2.log
