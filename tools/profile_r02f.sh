#!/bin/bash
# peer-memory exchange kernel vs NCCL path, N GPUs (N = the ranks of this box)
set -u
mkdir -p gpurun_out
N=${1:-2}
for q in 512 4096; do
CHECK_Q=$q timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/check_sharded.py 2>&1 | grep -E "sharded ==|Error|error" | tee -a gpurun_out/r02f_check_sharded_n$N.log
done
