#!/bin/bash
# decode attention: cache tiles requested before griddepcontrol.wait (SDAG_DECODE_EARLY_TILES=1, default) vs after
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_generate_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02z_pytest.log
out=gpurun_out/r02z_decode_early_tiles.log
: > $out
for rep in 1 2; do for e in 0 1; do
  echo "# SDAG_DECODE_EARLY_TILES=$e" >> $out
  SDAG_DECODE_EARLY_TILES=$e timeout 300 python tools/gen_bench.py --cases cfg2,cfg5 --new 64 >> $out 2>gpurun_out/r02z_err.log || tail -3 gpurun_out/r02z_err.log >> $out
done; done
for e in 0 1; do echo "# B=4 SDAG_DECODE_EARLY_TILES=$e" >> $out; SDAG_DECODE_EARLY_TILES=$e timeout 300 python tools/gen_bench.py --cases cfg2 --batch 4 --new 32 >> $out 2>>gpurun_out/r02z_err.log; done
python - <<'P'
import json
for l in open('gpurun_out/r02z_decode_early_tiles.log'):
    if l.startswith('#'): print(l.strip()); continue
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('   ', d['case'], 'B', d.get('batch', 1), round(d['decode_ms_per_token'], 3), 'ms/token')
P
