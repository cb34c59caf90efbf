"""bench.py's reference arm (the CPU leg: OpenMP port of the reference's loops, the one place besides tests/ and smoke()
that may execute oracle/) runs without a GPU; its JSON line must carry the contract's keys."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must use every host thread regardless
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--nx", "256", "--ny", "256",
                        "--steps", "2", "--warmup", "1", "--cpu-sample-rows", "64"], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "Mcell-steps/s" and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)), "the CPU arm must not inherit OMP_NUM_THREADS=1"
    assert d["steps"] == 2 and d["warmup"] == 1, "same steps / warm-up as asked for"
    assert d["cpu_baseline"]["sample_fraction"] == 64 / 256
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_ranks_other_than_zero_stay_silent_in_the_reference_arm():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--nx", "256", "--ny", "256"],
                       capture_output=True, text=True, timeout=60, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
