#!/bin/bash
mkdir -p gpurun_out
PYTHONPATH=. timeout 60 python tools/overlap_probe.py 8192 1 > gpurun_out/r1_overlap_probe_zc.json 2> gpurun_out/r1_overlap_probe_zc.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r1_overlap_probe_zc.json'))
for k,v in d.items(): print(k, round(v['all_done_ms'],1))
PY
tail -3 gpurun_out/r1_overlap_probe_zc.err
timeout 100 python bench.py --nx 8192 --ny 8192 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_8192_pipe3.json 2> gpurun_out/r1_bench_8192_pipe3.err
python -c "
import json
d=json.load(open('gpurun_out/r1_bench_8192_pipe3.json')); print(d['value'], d['e2e'])
"; tail -3 gpurun_out/r1_bench_8192_pipe3.err
