#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k skinny -x > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02n_pytest.log
out=gpurun_out/r02n_skinny_sweep.log
: > $out
run() { timeout 120 python tools/skinny_sweep.py ${B:-1} "$1" >> $out 2>&1 || echo "failed $*" >> $out; }
SDAG_SKINNY_DEBUG=1 run qkv,o,gate_up,down,head
B=8 run qkv,o,gate_up,down,head
SDAG_SKINNY_SMEM_KB=224 run qkv,o,down
SDAG_SKINNY_SMEM_KB=110 run gate_up,head
for sk in 2 4 8; do SDAG_SKINNY_SK=$sk run o; SDAG_SKINNY_SK=$sk run down;  SDAG_SKINNY_SK=$sk run qkv; done
grep -v "^llm_skinny: B=" $out | grep -v Traceback | grep -v "^  " ; grep "^llm_skinny: B=" $out | sort | uniq -c
for B in 1 4 8; do timeout 300 python tools/gemv_bench.py $B 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['layer'], 'B', d['B'], 'skinny', d.get('skinny_us'), 'gemv', d.get('gemv_us'), 'cublas', d.get('cublas_us'), 'err', d.get('skinny_max_abs_err'))
"; done
