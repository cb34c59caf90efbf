#!/bin/bash
# compute-sanitizer over a slice of the GPU suite ("new kernels must be clean").  Run on the GPU box:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'            (single-GPU tools)
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/sanitize.sh slabs'   (memcheck over a 2-rank slab run, all processes)
# memcheck: out-of-bounds / misaligned accesses (the reference overruns its face arrays, SURVEY App. A Q4 — this code
# must not); initcheck: reads of uninitialised device memory; racecheck: shared-memory hazards of the sweep kernels
# (mbarrier ring); synccheck: barrier misuse.  Logs: gpurun_out/sanitizer_*.log (copied to profiles/ by hand).
mkdir -p gpurun_out
if [ "$1" = "slabs" ]; then
  for tool in memcheck racecheck; do
    echo "== compute-sanitizer --tool $tool, 2 ranks (tests/mgpu_worker.py, IFX_MGPU_QUICK=1)"
    IFX_MGPU_QUICK=1 timeout 420 compute-sanitizer --tool $tool --target-processes all --error-exitcode 9 \
      --log-file gpurun_out/sanitizer_slabs_${tool}.%p.log \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_worker.py \
      > gpurun_out/sanitizer_slabs_${tool}_run.log 2>&1
    echo "exit $?"; tail -2 gpurun_out/sanitizer_slabs_${tool}_run.log
    grep -h "ERROR SUMMARY" gpurun_out/sanitizer_slabs_${tool}.*.log | sort | uniq -c
  done
  exit 0
fi
SEL='shipped_case_20_steps or full_step_cylinder or red_black or multigrid_matches or line_relaxation or probes_and_forces or restart_is_bit_identical'
for tool in memcheck initcheck racecheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 330 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer_$tool.log \
    python -m pytest tests -q -m gpu -p no:cacheprovider -k "$SEL" -x > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  echo "exit $?"; tail -2 gpurun_out/sanitizer_${tool}_pytest.log; grep "ERROR SUMMARY" gpurun_out/sanitizer_$tool.log | tail -2
done
