#!/bin/bash
# iteration on the CTA-pair attention kernel: parity ladder, tests, isolated timings, one ncu capture
set -u
mkdir -p gpurun_out
for s in 3 4; do
  echo "== step $s"; timeout 120 python tools/pair_dev.py $s 2>&1 | tail -12
done > gpurun_out/r02c_pair_dev.log 2>&1
cat gpurun_out/r02c_pair_dev.log
timeout 600 python -m pytest tests/test_sdag_pair_gpu.py -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02c_pytest.log
SDAG_ATTN_PAIR=1 timeout 300 python tools/attn_bench.py --cases cfg2,cfg2_causal,cfg5_K2,causal8k --iters 20 2>&1 | tee gpurun_out/r02c_attn_pair.log
SDAG_ATTN_PAIR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sdag_prefill_pair_kernel -s 4 -c 1 -o gpurun_out/r02c_prof_pair_causal8k -f python tools/attn_bench.py --cases causal8k --iters 3 > gpurun_out/r02c_ncu.log 2>&1; echo "ncu rc=$?"
