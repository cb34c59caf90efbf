#!/bin/bash
# short GPU call: CLI case tests + the pipelined end-to-end measurement on a reduced grid
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_cli_case.py -q -m gpu -p no:cacheprovider > gpurun_out/r1_cli_case_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r1_cli_case_tests.log
tail -30 gpurun_out/r1_cli_case_tests.log
timeout 200 python bench.py --nx 8192 --ny 8192 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_8192_pipe.json 2> gpurun_out/r1_bench_8192_pipe.err
echo "bench exit $?"; cat gpurun_out/r1_bench_8192_pipe.json; tail -5 gpurun_out/r1_bench_8192_pipe.err
timeout 100 python bench.py --nx 8192 --ny 8192 --steps 5 --warmup 3 --no-cpu-baseline --e2e-handles 1 > gpurun_out/r1_bench_8192_serial.json 2> gpurun_out/r1_bench_8192_serial.err
echo "bench exit $?"; python -c "
import json
for f in ('pipe','serial'):
    d=json.load(open('gpurun_out/r1_bench_8192_%s.json'%f)); print(f, d['value'], d['e2e'])
"
