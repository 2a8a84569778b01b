"""bench.py contract checks that need no GPU: the reference arm (CPU oracle) prints ONE JSON line with the agreed keys, also
under a torchrun-style environment (rank 0 works with all host cores, other ranks exit silently)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, gpus=1):
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--gpus", str(gpus), "--steps", "2",
           "--warmup", "1"]
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "Mparticles/s P2G+G2P" and j["unit"] == "Mparticles/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 2 and j["warmup"] == 1
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]


def test_reference_arm_under_torchrun_env():
    # torchrun pins OMP_NUM_THREADS=1: rank 0 must take the host cores back; the other ranks print nothing
    r0 = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"}, gpus=2)
    assert r0.returncode == 0, r0.stderr[-2000:]
    j = json.loads(r0.stdout.strip())
    assert j["n_gpus"] == 2 and j["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert "2 copies end to end" in j["config"]["workload"] and j["config"]["n_gpus"] == 2 and j["scaling"] == "weak"
    r1 = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1", "OMP_NUM_THREADS": "1"}, gpus=2)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
