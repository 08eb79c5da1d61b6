#!/bin/bash
# Round-1 refresh of the ncu evidence after the CTA-pair kernels (run under gpurun, 1 GPU).  Each ncu pass follows a
# plain run of the same command; numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r01b.json 2> gpurun_out/bench_r01b.err
echo "bench rc=$?"
CMD_A="python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline"
$CMD_A > gpurun_out/plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/launches_sdag.csv $CMD_A > gpurun_out/ncu_a.log 2>&1
echo "launch list A rc=$?"
CMD_B="python bench.py --steps 2 --warmup 1 --layers 1 --corpus-rows 2000000 --no-cpu-baseline"
$CMD_B > gpurun_out/plain_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ip_|tighten|rescore|exact_scan|cast_rows|topk' -c 200 --csv --log-file gpurun_out/launches_ip.csv $CMD_B > gpurun_out/ncu_b.log 2>&1
echo "launch list B rc=$?"
# last (largest) epoch of the first search over a 4M-row corpus: 13 GEMM launches per search
CMD_D="python bench.py --steps 1 --warmup 1 --layers 1 --corpus-rows 4000000 --no-cpu-baseline"
$CMD_D > gpurun_out/plain_d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ip_gemm_filter_kernel_2cta -s 12 -c 1 -o gpurun_out/prof_ip_gemm_2cta -f $CMD_D > gpurun_out/ncu_d.log 2>&1
echo "full capture D rc=$?"
CMD_H="python tools/gemm_bench.py"
$CMD_H > gpurun_out/plain_h.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'llm_gemm_kernel<1' -s 3 -c 1 -o gpurun_out/prof_llm_gemm_swiglu -f $CMD_H > gpurun_out/ncu_h.log 2>&1
echo "full capture H rc=$?"
python tools/gemm_bench.py --sustain > gpurun_out/gemm_f.log 2>&1
python tools/gen_bench.py > gpurun_out/gen_f.log 2>&1
for q in 1 32 256 1024 4096; do python tools/ip_bench.py --rows 10000000 --queries $q 2>&1 | tail -1; done > gpurun_out/ip_f.log
tail -3 gpurun_out/gen_f.log; cat gpurun_out/gemm_f.log gpurun_out/ip_f.log
ls -la gpurun_out | head -40
