"""bench.py contract checks that need no GPU: the reference arm (the oracle on the host cores) prints ONE JSON line with the keys the
driver reads, in the same metric / unit / config vocabulary as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS=os.environ.get("OMP_NUM_THREADS", "4"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("diffusion steps/sec") and d["value"] > 0 and d["steps"] == 1
    assert d["scaling"] in ("weak", "strong") and d["n_gpus"] == 1
    # the unmodified reference (vendored by oracle/build_ref.py into the git-ignored oracle/_ref) when present, else the oracle port
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "model", "networks.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port"), d["cpu_baseline"]
    assert d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"] and "16 images" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])


def test_training_reference_arm_prints_one_json_line():
    """--workload train --impl reference: the unmodified reference's optimize_parameters arithmetic on a bounded sample of the batch."""
    env = dict(os.environ, OMP_NUM_THREADS=os.environ.get("OMP_NUM_THREADS", "8"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "train", "--gpus", "1", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "steps/s" and d["metric"].startswith("training steps/sec") and d["value"] > 0
    assert d["config"]["global_batch"] == 64 and "scaled by 32" in d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0
