#!/bin/bash
# decode step at 3 / 4 prompts on the fused llm_gemv step (6 launches per layer) instead of cuBLAS + glue (10): warp-MMA vs FMA form
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r02v_gen_fused_b4.log
: > $out
for B in 3 4; do
  for cfg in "2 2" "4 2" "4 4"; do
    set -- $cfg
    echo "# B=$B SDAG_FUSED_GEMV_MAX_B=$1 SDAG_GEMV_FMA_MAX_B=$2" >> $out
    SDAG_FUSED_GEMV_MAX_B=$1 SDAG_GEMV_FMA_MAX_B=$2 timeout 300 python tools/gen_bench.py --cases cfg2 --batch $B --new 32 >> $out 2>gpurun_out/r02v_err.log || { echo "FAILED" >> $out; tail -5 gpurun_out/r02v_err.log >> $out; }
  done
done
python - <<'P'
import json
for l in open('gpurun_out/r02v_gen_fused_b4.log'):
    if l.startswith('#'): print(l.strip()); continue
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('   ', d['batch'], round(d['decode_ms_per_token'], 3), 'ms/token', round(d['decode_hbm_gbs']), 'GB/s')
P
