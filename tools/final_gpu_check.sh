mkdir -p gpurun_out
timeout 800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/r2_tests_final.log
timeout 600 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<'PY'
import json
t=open('gpurun_out/r2_bench_final.json').read(); d=json.loads(t[t.index('{'):])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity_check']['bit_exact'], 'refcuda', d['ref_cuda_baseline'].get('value'), 'cpu', d['cpu_baseline']['value'], 'frac', d['roofline']['frac'], d['roofline']['predictor']['frac'], d['clocks'])
for c in d.get('configs', []):
    print({k: c.get(k) for k in ('name','error','ms_per_step','value','poisson_iterations','poisson_converged','poisson_residual')}, (c.get('parity_check') or {}).get('bit_exact'))
PY
tail -2 gpurun_out/r2_bench_final.err
