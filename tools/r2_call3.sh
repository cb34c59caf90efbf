mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -15 | tee gpurun_out/r2_call3_tests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_v4 --launch-skip 30 --launch-count 1 -f -o gpurun_out/r2_ppe_mask python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r2_call3_ncu.log 2>&1
tail -3 gpurun_out/r2_call3_ncu.log
timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 4 > gpurun_out/r2_call3_bench.json 2> gpurun_out/r2_call3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_call3_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['poisson']['ms_per_launch'], d['clocks'])
PY
