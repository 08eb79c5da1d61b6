#!/bin/bash
# Round 2, final single-GPU pass (after the epilogue-warp and encoder changes): the parity suite and the bench line of the final tree
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02r_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r02r_pytest.log; grep -E "config 1 greedy|bf16 engine vs|top-5 overlap" gpurun_out/r02r_pytest.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r02r_bench_n1.json 2> gpurun_out/r02r_bench_n1.err
echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02r_bench_ref.json 2> gpurun_out/r02r_bench_ref.err
echo "bench reference rc=$?"
python -c "
import json
b=json.load(open('gpurun_out/r02r_bench_n1.json')); r=b['roofline']
print(round(b['value']), 'tok/s e2e', round(b['e2e']['value']), 'ms', round(b['ms_per_step'],2), {k:(round(v['launch_ms'],3), round(v['achieved'])) for k,v in r['projections'].items()})
print('attn', r['sdag_attention']['launch_ms'], r['sdag_attention']['frac'], 'dense', r['unmasked_tiles']['achieved'])
print('hf', b['same_gpu_hf']['sdpa_tok_s'], b['same_gpu_hf']['eager_tok_s'], 'cfg5', b['config5']['prefill_tok_s'], b['config5']['decode_ms_per_token'])
print('cfg4', b['config4']['seconds'], 'cfg1', b['config1']['seconds'], 'qps', b['secondary']['value'], b['secondary']['parity_check'])
r=json.load(open('gpurun_out/r02r_bench_ref.json')); print('ref', r['value'], r['ms_per_step'], r['cpu_baseline']['mask_builder'])
"
