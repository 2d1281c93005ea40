"""bench.py contract that can be checked without a GPU: the reference arm prints ONE JSON line with the agreed keys
(the driver launches it on every rank; ranks other than 0 print nothing and exit 0), and the product arm refuses to
run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"]


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(None, "--impl", "reference", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert not [k for k in KEYS if k not in d]
    assert d["impl"] == "reference" and d["unit"] == "Mpoints/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    r = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == ""
