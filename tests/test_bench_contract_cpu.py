"""CPU: the reference arm of bench.py (the reference's torch path, or the C port of it, timed on the host cores)
prints one JSON line with the keys the driver reads, and describes the same job as the native arm of the same
command line.  Uses the small config-1 workload so that it runs in seconds without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("kind", ["torch", "port"])
def test_reference_arm_prints_contract_line(kind):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "1", "--ref-frames-per-step", "2", "--ref-kind", kind],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "voxel_updates_per_s" and d["unit"] == "voxel-updates/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 1
    assert d["cpu_baseline"]["kind"] == ("reference" if kind == "torch" else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same `config` the native arm prints for these arguments
    sys.path.insert(0, ROOT)
    import bench
    plan = bench.Plan(bench.parse_args(["--workload", "cfg1", "--steps", "1", "--warmup", "1"]), 1)
    assert d["config"] == plan.config()


def test_both_cpu_paths_count_the_same_updates():
    """The torch reference and its C port integrate the same frames of the shared plan: identical update counts."""
    sys.path.insert(0, ROOT)
    import bench
    plan = bench.Plan(bench.parse_args(["--workload", "cfg1", "--steps", "1", "--warmup", "0"]), 1)
    a, _ = bench.make_cpu_reference(plan, "torch", 2)
    b, _ = bench.make_cpu_reference(plan, "port", 2)
    for j in range(2):
        fr = bench.host_frame(plan, plan.position(0, j)[0])
        assert a.integrate(fr) == b.integrate(fr) > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
