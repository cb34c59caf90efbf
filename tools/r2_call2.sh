mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/r2_call2_tests.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 4 > gpurun_out/r2_call2_bench.json 2> gpurun_out/r2_call2_bench.err
tail -c 2500 gpurun_out/r2_call2_bench.json; tail -3 gpurun_out/r2_call2_bench.err
