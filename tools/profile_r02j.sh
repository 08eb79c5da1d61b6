#!/bin/bash
# early exact re-scoring under the last GEMM epoch: parity tests, then A/B on the 8-GPU shard size and on 10 M rows
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_retrieval_gpu.py tests/test_edges_gpu.py -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02j_pytest.log
for e in 1 0 1 0; do
  echo "== SDAG_IP_EARLY_RESCORE=$e"; SDAG_IP_EARLY_RESCORE=$e python tools/ip_bench.py --rows 1250000 --queries 4096
done 2>&1 | tee gpurun_out/r02j_ip_shard_ab.log
for e in 1 0; do
  echo "== SDAG_IP_EARLY_RESCORE=$e"; SDAG_IP_EARLY_RESCORE=$e python tools/ip_bench.py --rows 10000000 --queries 4096,32
done 2>&1 | tee gpurun_out/r02j_ip_10m_ab.log
