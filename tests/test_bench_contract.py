"""bench.py contract on the CPU: the reference arm (`--impl reference`, the oracle timed on the host cores) prints ONE JSON
line with the keys the driver reads, names the same workload as the CUDA arm, uses every host core even when the launcher
exports OMP_NUM_THREADS=1 (torchrun does), and ranks other than 0 exit 0 without printing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log2n", "12",
                           "--steps", "1", "--warmup", "1", *args], env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_json_line():
    r = _run({"OMP_NUM_THREADS": "1", "RANK": "0", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "encode_zt_apply_samples_per_s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["n_gpus"] == 2
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == (os.cpu_count() or 1)
    assert cb["threads"] >= 1 and "sample" in cb
    if (os.cpu_count() or 1) > 1:
        assert cb["threads"] > 1, "the launcher's OMP_NUM_THREADS=1 must not make the CPU arm single-threaded"
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.workload_name(12)          # the CUDA arm prints the same string


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert r.returncode == 0 and r.stdout.strip() == ""
