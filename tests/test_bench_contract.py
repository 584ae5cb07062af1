"""bench.py contract on CPU: stdout carries exactly one JSON line whatever libraries write to fd 1, and the reference arm
(`--impl reference`: the oracle port on the host cores, a bounded sample) prints the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stdout_guard_leaves_exactly_one_json_line():
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.guard_stdout(); "
            "os.write(1, b'NCCL version banner\\n'); print('library chatter'); bench.emit({'a': 1})") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"a": 1}\n'
    assert "NCCL version banner" in r.stderr and "library chatter" in r.stderr


def test_reference_arm_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"], capture_output=True, text=True,
                       cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"].startswith("patch-rays/sec") and j["unit"] == "patch-rays/s"
    assert j["higher_is_better"] is True and j["steps"] == 1 and j["warmup"] == 2 and j["value"] > 0 and j["ms_per_step"] > 0
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "512 patches" in cb["sample"] and "16 levels" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, cwd=ROOT, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
