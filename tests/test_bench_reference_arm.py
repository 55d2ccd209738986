"""CPU: `bench.py --impl reference` (the arm the driver runs next to the GPU arm) on a small configuration.
It times the reference's own push_ptcls + search_mesh source (oracle/_ref, or the oracle port where that
is not built) on the host cores and must print ONE JSON line with the contract's keys; under torchrun
only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, args=()):
    env = dict(os.environ)
    env.update(env_extra or {})
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--particles", "40000", "--cube-n", "8"] + list(args)
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["metric"] == "particle push+search steps/s" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["particles_per_gpu"] == 40000 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm is the reference's code only: the product library is not loaded by it
    assert "libpumipic_b200" not in json.dumps(d.get("native_so_loaded", []))


def test_reference_arm_under_torchrun_only_rank_zero_works():
    base = {"WORLD_SIZE": "2", "LOCAL_RANK": "1", "RANK": "1", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29987"}
    r = _run(base, ["--gpus", "2"])
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout[-500:], r.stderr[-500:])
