#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== poly"; SDAG_ATTN_PAIR=1 timeout 300 python tools/attn_bench.py --cases cfg2,cfg5_K2,causal8k --iters 20 2>&1 | tee gpurun_out/r02c_attn_pair.log
echo "== no poly"; SDAG_B200_LIB=$PWD/sparse-document-attention-rag-_b200/_build_dbg/libsdag_b200_nopoly.so SDAG_ATTN_PAIR=1 timeout 300 python tools/attn_bench.py --cases cfg2,cfg5_K2,causal8k --iters 20 2>&1 | tee gpurun_out/r02c_attn_pair_nopoly.log
