"""GPU twin of tests/test_distributed.py: with >= 2 visible GPUs, two NCCL ranks each contract
their block of m10 slices on their own device and one NCCL reduce gives the amplitude."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from jet_b200 import ContractionPlan, NetworkFile
from jet_b200.distributed import slice_range, reduce_amplitude
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
net = NetworkFile.load(os.path.join({root!r}, "data", "_ref", "m10.json"), np.complex64)
plan = ContractionPlan(net, "p7 s7 h4 m1 m2 I2".split(), device=local)
first, count = slice_range(plan.num_slices, world, rank)
plan.reset(); plan.run(first, count)
total = reduce_amplitude(plan.result().reshape(1), device=torch.device("cuda", local))
if rank == 0:
    print("RESULT", json.dumps([float(total[0].real), float(total[0].imag)]))
plan.close()
dist.destroy_process_group()
'''


def test_nccl_world2_sliced_amplitude(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29544", str(script)],
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")][0]
    re, im = json.loads(line[len("RESULT "):])
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "amplitudes.json")))["m10_s6_sum64_complex128"]
    want = complex(gold["re"], gold["im"])
    assert abs(complex(re, im) - want) / abs(want) < 1e-5
