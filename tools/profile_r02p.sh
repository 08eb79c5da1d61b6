#!/bin/bash
# 8 epilogue warps for short main loops (K <= 2048): A/B on the encoder's projections, parity, encoder throughput
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_encoder_gpu.py tests/test_config1_gpu.py -m gpu -q -x > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02p_pytest.log
for ew in 4 8; do echo "== SDAG_GEMM_EW=$ew"; SDAG_GEMM_EW=$ew python tools/gemm_bench.py --encoder --tokens 40960 --sustain 2>&1 | tail -8; done > gpurun_out/r02p_gemm_enc_ab.log 2>&1
cat gpurun_out/r02p_gemm_enc_ab.log
for ew in 4 8; do echo "== SDAG_GEMM_EW=$ew"; SDAG_GEMM_EW=$ew timeout 300 python tools/encode_bench.py 2>&1 | tail -2; done > gpurun_out/r02p_encode_ab.log 2>&1
cat gpurun_out/r02p_encode_ab.log
