"""The reference arm of bench.py (the CPU port timed on the host cores) runs without a GPU: check the JSON contract of
its line, and that non-zero ranks of a torchrun launch stay silent and exit 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                           "--config", "small"],   # same code path as C3, sized for the CPU suite
                          capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)


def test_reference_arm_line():
    p = _run({})
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "d=2048 BSGS CKKS matvecs/s" and line["unit"] == "matvecs/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["config"]["workload"].startswith("SMALL: r,k,v projections of one RWKV-7 block = 3 x (64x64 BSGS mat-vec")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "rotation" in cb["sample"]
    # the sampled steps are anchored by one WHOLE un-hoisted mat-vec run in the same process, and they agree with it
    whole = line["whole_matvec"]
    assert whole["seconds"] > 0 and 0.2 < line["value"] / whole["value"] < 5.0
    # a step's duration is real: steps x ms_per_step fits inside the run (the driver checks the same against its own clock)
    assert line["steps"] * line["ms_per_step"] * 1e-3 < line["wall_s"]
    assert line["e2e"] == {"value": line["value"], "unit": "matvecs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    p = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""
