"""bench.py's JSON contract on the CPU side: the reference arm (the reference's own TaskBasedContractor through
oracle/_ref, the one place besides tests where bench.py executes the checker) prints one line with the keys the
driver reads.  The GPU arm is exercised on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libjetref.so")):
        pytest.skip("oracle/_ref not built")
    if not os.path.exists(os.path.join(ROOT, "data", "_ref", "m12.json")):
        pytest.skip("data/_ref missing")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--workload", "sycamore53_m12_s9"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "slices/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"] == "sycamore53_m12_s9"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0


def test_bench_workloads_resolve():
    sys.path.insert(0, ROOT)
    import bench
    if not os.path.exists(os.path.join(ROOT, "data", "_ref", "m12.json")):
        pytest.skip("data/_ref missing")
    net, sliced, dt, _ = bench.load_network("sycamore53_m12_s9")
    assert len(net.tensors) == 410 and len(net.path) == 409 and len(sliced) == 9 and dt == "complex64"
    net, sliced, dt, _ = bench.load_network("sycamore53_m20_synth")
    assert len(net.tensors) == 870 and len(sliced) == 58
    net, sliced, dt, _ = bench.load_network("sycamore53_m20")  # the default workload
    assert len(net.tensors) == 860 and len(sliced) == 22 and dt == "complex64"
