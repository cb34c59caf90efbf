#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r2_tests_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err ) 2>&1 | grep real
python - <<'PY'
import json
t=open('gpurun_out/r2_bench_final.json').read(); d=json.loads(t[t.index('{'):])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity_check']['bit_exact'], 'refcuda', d['ref_cuda_baseline'].get('value'), 'cpu', d['cpu_baseline']['value'], 'frac', d['roofline']['frac'], d['roofline']['predictor']['frac'])
for c in d.get('configs', []):
    print({k: c.get(k) for k in ('name','error','ms_per_step','value','poisson_iterations','poisson_converged','gpu_launches_per_step')})
PY
tail -2 gpurun_out/r2_bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --no-cpu-baseline --no-e2e --no-parity-check --no-ref-cuda --no-secondary --steps 1 --warmup 1 > gpurun_out/r2_launches_run.log 2>&1
wc -l gpurun_out/r2_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_v4 --launch-skip 30 --launch-count 1 -f -o gpurun_out/r2_final_ppe python bench.py --no-cpu-baseline --no-e2e --no-parity-check --no-ref-cuda --no-secondary --steps 1 --warmup 1 > gpurun_out/r2_final_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_sweep_v4 --launch-skip 10 --launch-count 1 -f -o gpurun_out/r2_final_ad python bench.py --no-cpu-baseline --no-e2e --no-parity-check --no-ref-cuda --no-secondary --steps 1 --warmup 1 > gpurun_out/r2_final_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
