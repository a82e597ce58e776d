"""bench.py's CPU arm (--impl reference) end to end on a tiny workload: it must print ONE JSON line with the keys the
driver reads, and every oracle run must happen in a bounded subprocess (no torch.set_num_threads in the parent)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--dz", "24", "--batch", "16"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "QP/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    # oracle/_ref (the unmodified reference, pip-installed by __graft_entry__.build()) when present, else the port
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "QP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the SAME config object as the GPU arm prints (one builder function): the driver's same_config check
    assert set(d["config"]) == {"workload", "global_batch", "dz", "n_eq", "admm_iter", "parallelism", "l2"}
    assert d["config"]["dz"] == 24 and d["config"]["global_batch"] == 16


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--dz", "24", "--batch", "16"], capture_output=True, text=True, timeout=300, env=env,
                       cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""
