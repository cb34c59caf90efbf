mkdir -p gpurun_out
for v in default forceslab; do
  if [ $v != default ]; then export IFX_LIBRARY=$PWD/tools/_bin/lib_$v.so; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_v4 --launch-skip 30 --launch-count 1 -f -o gpurun_out/r2_emu8_$v python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --emulate-slab-of 8 > gpurun_out/r2_call11_ncu_$v.log 2>&1
  tail -1 gpurun_out/r2_call11_ncu_$v.log
  timeout 600 ncu --set full --clock-control none -k regex:k_sweep_v4 --launch-skip 10 --launch-count 1 -f -o gpurun_out/r2_emu8_ad_$v python bench.py --no-cpu-baseline --no-e2e --steps 1 --warmup 1 --emulate-slab-of 8 > gpurun_out/r2_call11_ncu_ad_$v.log 2>&1
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --emulate-slab-of 8 > gpurun_out/r2_call11_$v.json 2> gpurun_out/r2_call11_$v.err
  python - <<PY
import json
t=open('gpurun_out/r2_call11_$v.json').read()
d=json.loads(t[t.index('{'):])
r=d['roofline']
print('$v', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'dev', round(d['device_ms_per_step'],2), 'ad/launch', round(r['ms_per_launch'],4), 'ppe/launch', round(r['poisson']['ms_per_launch'],4), d['clocks']['sm_mhz'])
PY
done
