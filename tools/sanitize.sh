#!/bin/bash
# compute-sanitizer over a slice of the GPU suite ("new kernels must be clean").  Run on the GPU box:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'            (single-GPU tools)
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/sanitize.sh slabs'   (memcheck over a 2-rank slab run, all processes)
# memcheck: out-of-bounds / misaligned accesses (the reference overruns its face arrays, SURVEY App. A Q4 — this code
# must not); initcheck: reads of uninitialised device memory; racecheck: shared-memory hazards of the sweep kernels
# (mbarrier ring); synccheck: barrier misuse.  Logs: gpurun_out/sanitizer_*.log (copied to profiles/ by hand).
mkdir -p gpurun_out
if [ "$1" = "slabs" ]; then
  # every rank runs under its own sanitizer (no launcher in between, so the tool certainly sees the worker's kernels)
  for tool in memcheck racecheck; do
    echo "== compute-sanitizer --tool $tool, 2 ranks (tests/mgpu_worker.py, IFX_MGPU_QUICK=1), one sanitizer per rank"
    for r in 0 1; do
      IFX_MGPU_QUICK=1 RANK=$r LOCAL_RANK=$r WORLD_SIZE=2 MASTER_ADDR=127.0.0.1 MASTER_PORT=29541 \
        timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer_slabs_${tool}_rank$r.log \
        python tests/mgpu_worker.py > gpurun_out/sanitizer_slabs_${tool}_rank${r}_run.log 2>&1 &
    done
    wait
    for r in 0 1; do
      tail -1 gpurun_out/sanitizer_slabs_${tool}_rank${r}_run.log
      grep -h "SUMMARY" gpurun_out/sanitizer_slabs_${tool}_rank$r.log
      grep -c "k_sweep_v4\|k_halo\|k_gc" gpurun_out/sanitizer_slabs_${tool}_rank$r.log
    done
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
