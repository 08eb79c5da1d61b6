#!/bin/bash
# config 1 / config 4 stage seconds: 8 vs 4 epilogue warps (is the slower generate_carg of the last pass the code or the box?)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LITE="--no-secondary --no-cpu-baseline --no-hf --no-config5 --steps 3 --warmup 3 --config4-queries 16"
for ew in 8 4 8 4; do
  SDAG_GEMM_EW=$ew python bench.py $LITE > gpurun_out/r02s_bench_ew$ew.json 2> gpurun_out/r02s_err.log || tail -5 gpurun_out/r02s_err.log
  python -c "
import json
b=json.load(open('gpurun_out/r02s_bench_ew$ew.json'))
print('EW=$ew', 'cfg1', round(b['config1']['seconds'],3), b['config1']['stage_seconds'], 'cfg4', round(b['config4']['seconds'],2), b['config4']['per_strategy']['closest_to_centroid']['stage_seconds'], 'encode', round(b['config4']['encode_corpus_seconds'],2))
"
done
