#!/bin/bash
# ncu on the encoder's K = 1024 projections (llm_gemm_residual N = 1024, llm_gemm GELU N = 4096) at 40960 token rows
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/gemm_bench.py --encoder --tokens 40960 --only "enc o+bias+res" > gpurun_out/r02o_plain1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:llm_gemm -s 4 -c 1 -o gpurun_out/r02o_prof_enc_o python tools/gemm_bench.py --encoder --tokens 40960 --only "enc o+bias+res" > gpurun_out/r02o_ncu1.log 2>&1
echo "ncu o rc=$?"
python tools/gemm_bench.py --encoder --tokens 40960 --only "enc ffn1+gelu" > gpurun_out/r02o_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:llm_gemm -s 4 -c 1 -o gpurun_out/r02o_prof_enc_ffn1 python tools/gemm_bench.py --encoder --tokens 40960 --only "enc ffn1+gelu" > gpurun_out/r02o_ncu2.log 2>&1
echo "ncu ffn1 rc=$?"
ls -la gpurun_out | grep r02o
