#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out, in one go.
#   gpurun --timeout 900 -- 'bash tools/r2_first_gpu_call.sh'
# Unverified on a GPU at the time of writing (all bit-identical to the oracle on the CPU through tests/shim):
#   bilinear prolongation of the line-smoothed cycle, odd-count coarsening, L1 prefetch hints + chunked back substitution
#   of the line kernels, CUDA-graph replay of the coarse cycle (ifx_options.use_graphs), body-inside-grid check,
#   the end-to-end pipeline at 16384 x 16384 (measured at 8192 x 8192 only).
mkdir -p gpurun_out
echo "== parity suites" | tee gpurun_out/r2_first.log
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15 | tee -a gpurun_out/r2_first.log
echo "== multigrid / SOR on the cavity (plain launches, then graph replay)" | tee -a gpurun_out/r2_first.log
PYTHONPATH=. timeout 120 python tools/mg_bench.py | tee gpurun_out/r2_mg_bench.jsonl
IFX_MG_BENCH_GRAPHS=1 PYTHONPATH=. timeout 120 python tools/mg_bench.py | tee gpurun_out/r2_mg_bench_graphs.jsonl
echo "== line-smoothed multigrid on stretched grids" | tee -a gpurun_out/r2_first.log
PYTHONPATH=. timeout 120 python tools/line_mg_bench.py 1024 5 | tee gpurun_out/r2_line_mg_1024.jsonl
PYTHONPATH=. timeout 200 python tools/line_mg_bench.py 4096 5 | tee gpurun_out/r2_line_mg_4096.jsonl
echo "== default bench (16384 x 16384, pipelined end-to-end)" | tee -a gpurun_out/r2_first.log
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 1500 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
echo "== sanitizers"; bash tools/sanitize.sh 2>&1 | tee -a gpurun_out/r2_first.log
