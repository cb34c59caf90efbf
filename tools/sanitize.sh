#!/bin/bash
# compute-sanitizer over a small slice of the GPU suite ("new kernels must be clean").  Run on the GPU box:
#   gpurun --timeout 900 -- 'bash tools/sanitize.sh'
# memcheck: out-of-bounds / misaligned accesses (the reference overruns its face arrays, SURVEY App. A Q4 — this
# code must not); initcheck: reads of uninitialised device memory; racecheck: shared-memory hazards of the sweep kernels.
mkdir -p gpurun_out
SEL='shipped_case_20_steps or full_step_cylinder or multigrid_matches or line_relaxation or probes_and_forces'
for tool in memcheck initcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 280 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitizer_$tool.log \
    python -m pytest tests -q -m gpu -p no:cacheprovider -k "$SEL" -x > gpurun_out/sanitizer_${tool}_pytest.log 2>&1
  echo "exit $?"; tail -3 gpurun_out/sanitizer_${tool}_pytest.log; grep -c "ERROR SUMMARY" gpurun_out/sanitizer_$tool.log; grep "ERROR SUMMARY" gpurun_out/sanitizer_$tool.log | tail -2
done
