#!/bin/bash
# Round 2, second GPU pass (1 GPU): parity suite, isolated attention timings, same-box A/B of the o / down projections
# (this repo's GEMM with the residual epilogue vs cuBLAS + stand-alone add), shard-sized retrieval timings, then ncu:
# DRAM counters of the stand-alone top-k kernels and full captures of the attention kernel (config 2) and the wide GEMM.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r02b_pytest.log
python tools/attn_bench.py --cases cfg2,cfg2_causal,cfg5_K2,causal8k --iters 20 > gpurun_out/r02b_attn.log 2>&1; cat gpurun_out/r02b_attn.log
LITE="--no-secondary --no-cpu-baseline --no-hf --no-config5 --no-pipeline --steps 10 --warmup 3"
python bench.py $LITE > gpurun_out/r02b_bench_ours.json 2> gpurun_out/r02b_bench_ours.err; echo "bench ours rc=$?"
SDAG_CUBLAS_PROJ=o,down python bench.py $LITE > gpurun_out/r02b_bench_cublas_o_down.json 2> gpurun_out/r02b_bench_cublas.err; echo "bench cublas o,down rc=$?"
python bench.py $LITE > gpurun_out/r02b_bench_ours2.json 2>> gpurun_out/r02b_bench_ours.err; echo "bench ours(2) rc=$?"
python -c "
import json
for n in ('ours','cublas_o_down','ours2'):
    b=json.load(open(f'gpurun_out/r02b_bench_{n}.json')); r=b['roofline']
    print(n, round(b['value']), 'tok/s', round(b['ms_per_step'],2), 'ms', {k:(round(v['launch_ms'],3), round(v['achieved'])) for k,v in r['projections'].items()}, 'attn', round(r['sdag_attention']['launch_ms'],4), 'dense', round(r['unmasked_tiles']['achieved']), b['clocks'])
"
python tools/gemm_bench.py --sustain > gpurun_out/r02b_gemm.log 2>&1; tail -6 gpurun_out/r02b_gemm.log
python tools/ip_bench.py --rows 1250000 --queries 4096,32 > gpurun_out/r02b_ip_shard.log 2>&1; cat gpurun_out/r02b_ip_shard.log
CMD_T="python tools/ip_bench.py --rows 1250000 --queries 4096"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__bytes_read.sum.per_second,dram__bytes_write.sum.per_second --clock-control none -k regex:'tighten|rescore|topk|exact_scan|cast_rows|init_state' -c 80 --csv --log-file gpurun_out/r02b_topk_dram.csv $CMD_T > gpurun_out/r02b_ncu_t.log 2>&1
echo "ncu topk rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sdag_prefill_kernel -s 4 -c 1 -o gpurun_out/r02b_prof_sdag_cfg2 -f python tools/attn_bench.py --cases cfg2 --iters 3 > gpurun_out/r02b_ncu_s.log 2>&1
echo "ncu sdag rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:llm_gemm_wide_kernel -s 14 -c 1 -o gpurun_out/r02b_prof_gemm_wide -f python tools/gemm_bench.py --only down+residual > gpurun_out/r02b_ncu_w.log 2>&1
echo "ncu wide rc=$?"
ls -la gpurun_out | tail -16
