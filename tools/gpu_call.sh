#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slabs.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3 | tee gpurun_out/r2_slabs_tests8.log
run() { # name, nproc, extra args
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-ref-cuda --no-secondary $3 > gpurun_out/r2_scale_$1.json 2> gpurun_out/r2_scale_$1.err
  python - <<PY
import json
try:
    t=open('gpurun_out/r2_scale_$1.json').read()
    d=json.loads(t[t.index('{'):])
    r=d['roofline']
    print('$1', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'dev', round(d['device_ms_per_step'],2), 'ad/launch', round(r['predictor']['ms_per_launch'],4), 'ppe/launch', round(r['ms_per_launch'],4), 'proj', round(r['projection_ms'],3), 'ib', round(r['iblank_ghost_cells_ms'],3), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['gpu_launches'], 'parity', (d.get('parity_check') or {}).get('bit_exact'), d.get('slab_parity'))
except Exception as e:
    print('$1', e); print(open('gpurun_out/r2_scale_$1.err').read()[-800:])
PY
}
run n8 8 ""
run n8_rows 8 "--slab-balance rows --no-parity-check"
run n4 4 ""
run n2 2 ""
run n1 1 ""
