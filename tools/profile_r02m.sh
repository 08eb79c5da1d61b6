#!/bin/bash
# llm_skinny_linear bring-up: parity test, then per-layer-shape streaming rate vs llm_gemv and cuBLAS at 1 / 4 / 8 prompts
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k skinny -x > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02m_pytest.log
: > gpurun_out/r02m_gemv_bench.log
for B in 1 4 8; do timeout 300 python tools/gemv_bench.py $B >> gpurun_out/r02m_gemv_bench.log 2>&1 || echo "bench B=$B failed" >> gpurun_out/r02m_gemv_bench.log; done
cat gpurun_out/r02m_gemv_bench.log
