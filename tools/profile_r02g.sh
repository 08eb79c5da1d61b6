#!/bin/bash
# Round 2: full parity suite on the final tree, A/B of the down projection on wide vs narrow tiles, encoder throughput + launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02g_pytest.log
LITE="--no-secondary --no-cpu-baseline --no-hf --no-config5 --no-pipeline --steps 10 --warmup 3"
python bench.py $LITE > gpurun_out/r02g_bench_wide.json 2> gpurun_out/r02g_bench.err; echo "bench (down on wide tiles) rc=$?"
SDAG_GEMM_WIDE=0 python bench.py $LITE > gpurun_out/r02g_bench_narrow.json 2>> gpurun_out/r02g_bench.err; echo "bench (down on narrow tiles) rc=$?"
python bench.py $LITE > gpurun_out/r02g_bench_wide2.json 2>> gpurun_out/r02g_bench.err
python -c "
import json
for n in ('wide','narrow','wide2'):
    b=json.load(open(f'gpurun_out/r02g_bench_{n}.json')); r=b['roofline']
    print(n, round(b['value']), 'tok/s', round(b['ms_per_step'],2), 'ms', {k:(round(v['launch_ms'],3), round(v['achieved'])) for k,v in r['projections'].items()}, b['clocks']['sm_mhz'])
"
python tools/encode_bench.py > gpurun_out/r02g_encode.log 2>&1; tail -1 gpurun_out/r02g_encode.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02g_launches_enc.csv python tools/encode_bench.py --docs 1024 > gpurun_out/r02g_ncu_enc.log 2>&1; echo "ncu enc rc=$?"
