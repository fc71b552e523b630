"""bench.py's contract, as far as a box without a GPU can check it: the reference arm (the CPU oracle port on a bounded sample)
prints exactly one JSON line on stdout with the fields the driver reads, and the GPU arm's line is assembled from the same
config block."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-images", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "extract_1080p_images_per_s" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "1920x1080" in d["config"]["workload"]


def test_gpu_arm_keeps_stdout_for_the_json_line():
    """The GPU arm routes file descriptor 1 to stderr while it runs (library banners such as NCCL's version line) and restores it
    for the one JSON line: checked on the source, the run itself needs a GPU."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "os.dup2(2, 1)" in src and "os.dup2(real_stdout, 1)" in src
    assert src.count("print(json.dumps(line))") == 2  # one per arm
