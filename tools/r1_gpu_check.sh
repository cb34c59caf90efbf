#!/bin/bash
# one short GPU call: the parity tests of the components added late in round 1 + a multigrid timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r1_late_gpu.txt 2>&1
timeout 150 python -m pytest tests/test_gpu_multigrid.py tests/test_gpu_diag.py -q -m gpu -p no:cacheprovider > gpurun_out/r1_late_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1_late_tests.log
tail -25 gpurun_out/r1_late_tests.log
timeout 100 python tools/mg_bench.py > gpurun_out/r1_mg_bench.jsonl 2> gpurun_out/r1_mg_bench.err
echo "mg_bench exit $?"
cat gpurun_out/r1_mg_bench.jsonl; tail -5 gpurun_out/r1_mg_bench.err
