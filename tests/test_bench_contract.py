"""CPU suite: the reference arm of bench.py (`--impl reference`) runs without a GPU - check that it prints exactly one
JSON line with every key the bench contract names, on a small sample so the suite stays fast.  The line's numbers are
not asserted (this container is not the GPU box); its shape is."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_CPU_SIDE = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_harness")) or os.path.exists(os.path.join(ROOT, "oracle", "gl_oracle"))


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cpu-sample-ngas", "4000"] + extra, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.strip()]


@pytest.mark.skipif(not HAVE_CPU_SIDE, reason="oracle binaries not built (run __graft_entry__.build())")
def test_reference_arm_prints_one_contract_line():
    lines = _run([])
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "interactions/s" and d["value"] > 0 and d["ms_per_step"] > 0
    # same workload string as our arm's line (BASELINE config 0), the bounded sample is named beside it
    assert d["config"]["workload"].startswith("BASELINE config 0") and "1600000 gas" in d["config"]["workload"]
    assert "4000" in d["config"]["sample"] and d["config"]["particles_per_instance"] == 4002
    assert d["steps"] == 1 and d["warmup"] == 1       # --steps / --warmup are honoured as given
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)


@pytest.mark.skipif(not HAVE_CPU_SIDE, reason="oracle binaries not built (run __graft_entry__.build())")
def test_reference_arm_other_ranks_exit_quietly():
    # under torchrun only rank 0 runs and prints the reference line
    assert _run(["--gpus", "2"], env=dict(RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")) == []
