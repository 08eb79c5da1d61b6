#!/bin/bash
# ncu evidence for round 1 (run under gpurun, 1 GPU).  Each ncu pass follows a plain run of the same command.
set -u
mkdir -p gpurun_out
python bench.py --carg > gpurun_out/bench_r01_f.json 2> gpurun_out/bench_r01_f.err
echo "bench rc=$?"
CMD_A="python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline"
$CMD_A > gpurun_out/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/launches_sdag.csv $CMD_A > gpurun_out/ncu_a.log 2>&1
echo "launch list A rc=$?"
CMD_B="python bench.py --steps 2 --warmup 1 --layers 1 --corpus-rows 2000000 --no-cpu-baseline"
$CMD_B > gpurun_out/plain_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ip_|tighten|rescore|exact_scan|cast_rows|topk' -c 200 --csv --log-file gpurun_out/launches_ip.csv $CMD_B > gpurun_out/ncu_b.log 2>&1
echo "launch list B rc=$?"
CMD_C="python tools/attn_bench.py --iters 2 --cases cfg2"
$CMD_C > gpurun_out/plain_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sdag_prefill_kernel -s 2 -c 1 -o gpurun_out/prof_sdag_prefill_cfg2 -f $CMD_C > gpurun_out/ncu_c.log 2>&1
echo "full capture C rc=$?"
CMD_E="python tools/attn_bench.py --iters 2 --cases causal8k"
$CMD_E > gpurun_out/plain_e.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sdag_prefill_kernel -s 2 -c 1 -o gpurun_out/prof_sdag_causal8k -f $CMD_E > gpurun_out/ncu_e.log 2>&1
echo "full capture E rc=$?"
CMD_G="python tools/attn_bench.py --iters 2 --decode --cases cfg5_K2"
$CMD_G > gpurun_out/plain_g.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:decode_split_kernel -s 2 -c 1 -o gpurun_out/prof_decode_cfg5 -f $CMD_G > gpurun_out/ncu_g.log 2>&1
echo "full capture G rc=$?"
# last (largest) epoch of the first search over a 4M-row corpus: 13 GEMM launches per search
CMD_D="python bench.py --steps 1 --warmup 1 --layers 1 --corpus-rows 4000000 --no-cpu-baseline"
$CMD_D > gpurun_out/plain_d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ip_gemm_filter_kernel -s 12 -c 1 -o gpurun_out/prof_ip_gemm -f $CMD_D > gpurun_out/ncu_d.log 2>&1
echo "full capture D rc=$?"
python tools/attn_bench.py --iters 20 > gpurun_out/attn_f.log 2>&1; python tools/attn_bench.py --iters 10 --decode --cases cfg2,cfg2_b1,cfg2_b32,cfg5_K2,causal8k 2>&1 | grep decode >> gpurun_out/attn_f.log
python tools/gen_bench.py > gpurun_out/gen_f.log 2>&1
tail -3 gpurun_out/gen_f.log; cat gpurun_out/attn_f.log
ls -la gpurun_out | head -40
