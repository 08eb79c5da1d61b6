#!/bin/bash
# Round 2, 8-GPU pass: sharded top-k == single-index top-k over NCCL (packed records, one all-gather), then the bench line at N = 8.
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py > gpurun_out/r02e8_check_sharded_n8.log 2>&1
echo "check_sharded rc=$?"; tail -2 gpurun_out/r02e8_check_sharded_n8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --balance-shards > gpurun_out/r02e8_bench_n8.json 2> gpurun_out/r02e8_bench_n8.err
echo "bench n8 rc=$?"; tail -c 400 gpurun_out/r02e8_bench_n8.err
python -c "
import json
b=json.loads([l for l in open('gpurun_out/r02e8_bench_n8.json') if l.startswith('{')][-1])
s=b['secondary']; print('tok/s', round(b['value']), 'qps', round(s['value']), 'ms', round(s['ms_per_step'],3), 'frac', round(s['roofline']['frac'],3), s['parity_check'])
print('config4', {k:b['config4'][k] for k in ('queries','seconds','queries_per_s')}, 'config1', b['config1']['queries_per_s'])
"
python -c "
import json
b=json.loads([l for l in open('gpurun_out/r02e8_bench_n8.json') if l.startswith('{')][-1]); s=b['secondary']
print('balanced+pipelined', round(s['value']), round(s['ms_per_step'],3), 'lockstep', s.get('lockstep'), 'rows', s['rows_per_rank'], 'local', s['rank_local_ms'])
print('equal', s.get('equal_shards'))
"
