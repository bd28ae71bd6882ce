"""bench.py's reference arm, as the driver launches it: one JSON line with the contract's keys from rank 0, silence and
exit 0 from every other rank (no GPU is involved: the arm times the unmodified HW3 binary on the host cores)."""
import json
import os
import subprocess
import sys

import pytest

from oracle import cpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(env_extra, *flags):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *flags], cwd=ROOT,
                          env=env, capture_output=True, text=True, timeout=900)


def test_other_ranks_of_the_reference_arm_do_no_work():
    r = _bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_prints_the_contract_line():
    if not cpu.have_ref("hw3"):
        pytest.skip("oracle/_ref not built on this box")
    r = _bench({"RANK": "0"}, "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "Mrays/sec" and line["unit"] == "Mrays/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("cfg2") and line["detail"]["rays_per_step"] == 5939386
    assert set(line["config"]) == {"workload", "split", "l2"} and line["config"]["split"] == "frame"   # what the own arm prints too
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_own_arm_refuses_to_run_without_a_device():
    """The product arm never falls back: without a GPU it says so and exits non-zero instead of timing the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)
    assert not any(l.startswith("{") for l in r.stdout.splitlines())
