mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x 2>&1 | tail -5 | tee gpurun_out/r2_call4_tests.log
for v in default min5; do
  if [ $v != default ]; then export IFX_LIBRARY=$PWD/tools/_bin/lib_$v.so; fi
  timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 4 > gpurun_out/r2_call4_bench_$v.json 2> gpurun_out/r2_call4_bench_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_call4_bench_$v.json'))
print('$v', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['poisson']['ms_per_launch'], d['clocks'])
PY
done
unset IFX_LIBRARY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_v4 --launch-skip 30 --launch-count 1 -f -o gpurun_out/r2_ppe_mask2 python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 > gpurun_out/r2_call4_ncu.log 2>&1
tail -2 gpurun_out/r2_call4_ncu.log
