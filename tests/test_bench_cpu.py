"""bench.py harness checks that need no GPU: the reference arm (the oracle on the host cores) prints the contract's JSON line,
sizes its OpenMP pool itself (torchrun exports OMP_NUM_THREADS=1) and runs the GPU arm's step shape."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_thread_count():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                          "--warmup", "1", "--labels", "2"], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "edges/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1), "the CPU arm must use every host core even under OMP_NUM_THREADS=1"
    assert d["cpu_baseline"]["kind"] == "port"
    assert d["config"]["labels_per_step"] == 2 and d["config"]["seeds_per_label"] == 1024
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0


def test_non_zero_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny"], env=env,
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
