#!/bin/bash
# decode step at 4 / 8 prompts: cuBLAS + glue vs llm_gemv (warp-MMA form) / llm_skinny_linear per linear layer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k skinny > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02l_pytest.log
out=gpurun_out/r02l_gen_batch.log
: > $out
for B in 4 8; do
  for G in "" "o=skinny" "o=skinny,down" "o=skinny,down=skinny" "qkv=skinny,o=skinny,gate_up=skinny,down=skinny,head=skinny" "o=skinny,down,head=skinny"; do
    SDAG_DECODE_GEMV="$G" timeout 300 python tools/gen_bench.py --cases cfg2 --batch $B --new 32 >> $out 2>gpurun_out/r02l_err.log || { echo "FAILED B=$B G=$G" >> $out; tail -5 gpurun_out/r02l_err.log >> $out; }
  done
done
python - <<'P'
import json
for l in open('gpurun_out/r02l_gen_batch.log'):
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print(d['batch'], repr(d['decode_gemv']), round(d['decode_ms_per_token'], 3), 'ms/token', round(d['decode_hbm_gbs']), 'GB/s')
P
