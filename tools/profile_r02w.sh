#!/bin/bash
# Final tree: ncu launch lists of the prefill step and of the encoder forward (refresh of profiles/r02_launches_*.md)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD_A="python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --no-hf --no-config5 --no-pipeline"
timeout 600 $CMD_A > gpurun_out/r02w_plain_a.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02w_launches_sdag.csv $CMD_A > gpurun_out/r02w_ncu_a.log 2>&1
echo "launch list prefill rc=$?"
CMD_E="python tools/encode_bench.py --docs 1024"
timeout 600 $CMD_E > gpurun_out/r02w_plain_e.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02w_launches_enc.csv $CMD_E > gpurun_out/r02w_ncu_e.log 2>&1
echo "launch list encoder rc=$?"
ls -la gpurun_out | grep r02w
