#!/bin/bash
mkdir -p gpurun_out
python tools/ref_cuda_baseline.py | tee gpurun_out/r2_ref_cuda_tool.json
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
t=open('gpurun_out/r2_bench_n1.json').read(); d=json.loads(t[t.index('{'):])
print(d['value'], d['e2e']['value'], d['parity_check']['bit_exact'], d.get('ref_cuda_baseline'))
PY
tail -3 gpurun_out/r2_bench_n1.err
