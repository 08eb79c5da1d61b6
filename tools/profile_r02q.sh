#!/bin/bash
# same-box A/B inside the prefill step: 4 vs 8 epilogue warps for the plain / residual / RoPE epilogues at K >= 4096
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02q_pytest.log
SDAG_GEMM_EW=8 timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_fullsize_gpu.py tests/test_generate_gpu.py -m gpu -q -x > gpurun_out/r02q_pytest_ew8.log 2>&1; echo "pytest (EW=8) rc=$?"; tail -3 gpurun_out/r02q_pytest_ew8.log
LITE="--no-secondary --no-cpu-baseline --no-hf --no-config5 --no-pipeline --steps 10 --warmup 3"
python bench.py $LITE > gpurun_out/r02q_bench_ew4.json 2> gpurun_out/r02q_bench.err; echo "bench default rc=$?"
SDAG_GEMM_EW=8 python bench.py $LITE > gpurun_out/r02q_bench_ew8.json 2>> gpurun_out/r02q_bench.err; echo "bench EW=8 rc=$?"
python bench.py $LITE > gpurun_out/r02q_bench_ew4b.json 2>> gpurun_out/r02q_bench.err; echo "bench default(2) rc=$?"
SDAG_GEMM_EW=8 python bench.py $LITE > gpurun_out/r02q_bench_ew8b.json 2>> gpurun_out/r02q_bench.err; echo "bench EW=8(2) rc=$?"
python -c "
import json
for n in ('ew4','ew8','ew4b','ew8b'):
    b=json.load(open(f'gpurun_out/r02q_bench_{n}.json')); r=b['roofline']
    print(n, round(b['value']), 'tok/s', round(b['ms_per_step'],2), 'ms', {k:(round(v['launch_ms'],3), round(v['achieved'])) for k,v in r['projections'].items()}, b['clocks']['sm_mhz'])
"
