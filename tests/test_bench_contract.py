"""bench.py contract checks that need no GPU: the CPU (reference) arm prints one well-formed JSON line, and the
static parts of the GPU arm's line (keys, algorithmic-byte constants) match DESIGN.md / SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun exports; the arm must override it
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "rays" in cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_bytes_match_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    # traverse 16, composite fwd 44, bwd 48 bytes per sample, 88 bytes per ray
    assert (bench.B_TRAVERSE, bench.B_FWD, bench.B_BWD, bench.B_RAY) == (16, 44, 48, 88)
    assert bench.METRIC.startswith("ray-samples/sec") and bench.RAYS_PER_GPU == 65536 and bench.GRID_RES == 128
