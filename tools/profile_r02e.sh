#!/bin/bash
# event trace of the CTA-pair attention kernel (debug build: -DSDAG_DEBUG_TIMING)
set -u
mkdir -p gpurun_out
SDAG_B200_LIB=$PWD/sparse-document-attention-rag-_b200/_build_dbg/libsdag_b200_dbg.so SDAG_ATTN_PAIR=1 timeout 120 python tools/substep_trace.py causal8k 120 > gpurun_out/r02e_trace_pair.log 2>&1
echo rc=$?; head -60 gpurun_out/r02e_trace_pair.log
