mkdir -p gpurun_out
echo "== parity suites" | tee gpurun_out/r2_first.log
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -15 | tee -a gpurun_out/r2_first.log
PYTHONPATH=. timeout 120 python tools/mg_bench.py | tee gpurun_out/r2_mg_bench.jsonl
IFX_MG_BENCH_GRAPHS=1 PYTHONPATH=. timeout 120 python tools/mg_bench.py | tee gpurun_out/r2_mg_bench_graphs.jsonl
PYTHONPATH=. timeout 120 python tools/line_mg_bench.py 1024 5 | tee gpurun_out/r2_line_mg_1024.jsonl
PYTHONPATH=. timeout 200 python tools/line_mg_bench.py 4096 5 | tee gpurun_out/r2_line_mg_4096.jsonl
timeout 600 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
tail -c 3000 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
