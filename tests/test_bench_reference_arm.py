"""bench.py --impl reference (the CPU arm the driver runs beside ours): prints ONE JSON line with the base contract's keys, runs no
GPU code and loads none of our .so files.  With the reference importable (the tree here, oracle/_ref on the GPU box) the arm is the
unmodified reference (kind "reference"); MM_BENCH_CPU=port forces the oracle port."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = {**os.environ, **env_extra, "CUDA_VISIBLE_DEVICES": ""}
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, env=env, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


@pytest.mark.parametrize("force_port", [False, True])
def test_reference_arm_line(force_port):
    from oracle import ref_harness
    d = _run({"MM_BENCH_CPU": "port"} if force_port else {})
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1
    want = "port" if (force_port or not ref_harness.available()) else "reference"
    assert cb["kind"] == want, cb
