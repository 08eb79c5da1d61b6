#!/bin/bash
# Round 2, first GPU pass (1 GPU): parity suite, the bench line, then the ncu launch list of the prefill step.
# Each ncu pass follows a plain run of the same command; numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02a_pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02a_bench.err
CMD_A="python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --no-hf --no-config5 --no-pipeline"
timeout 600 $CMD_A > gpurun_out/r02a_plain_a.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02a_launches_sdag.csv $CMD_A > gpurun_out/r02a_ncu_a.log 2>&1
echo "launch list A rc=$?"
python tools/gemm_bench.py --sustain > gpurun_out/r02a_gemm.log 2>&1; tail -12 gpurun_out/r02a_gemm.log
ls -la gpurun_out | tail -12
