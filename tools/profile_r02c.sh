#!/bin/bash
# bring-up of the CTA-pair attention kernel: each step in its own process under a timeout (a protocol bug traps after 2 s)
set -u
mkdir -p gpurun_out
for s in 1 2 3 4; do
  echo "== step $s"; timeout 120 python tools/pair_dev.py $s 2>&1 | tail -12
done > gpurun_out/r02c_pair_dev.log 2>&1
cat gpurun_out/r02c_pair_dev.log
timeout 600 python -m pytest tests/test_sdag_pair_gpu.py tests/test_config1_gpu.py -x -q -s > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02c_pytest.log
