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
from jet_b200 import ContractionPlan, MultiPlan, NetworkFile
from jet_b200.distributed import slice_range, reduce_amplitude, make_communicator, reduce_on_device
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
net = NetworkFile.load(os.path.join({root!r}, "data", "_ref", "m10.json"), np.complex64)
plan = ContractionPlan(net, "p7 s7 h4 m1 m2 I2".split(), device=local)
first, count = slice_range(plan.num_slices, world, rank)
plan.reset(); plan.run(first, count)
total = reduce_amplitude(plan.result().reshape(1), device=torch.device("cuda", local))
# the product path: NCCL inside libjetb200.so, device buffers in and out, on the plan's stream
comm = make_communicator(local)
multi = MultiPlan(net, "p7 s7 h4 m1 m2 I2".split(), lanes=2, device=local)
multi.run(first, count)
total2 = reduce_on_device(multi, comm, 0)
if rank == 0:
    print("RESULT", json.dumps([float(total[0].real), float(total[0].imag)]))
    print("RESULT2", json.dumps([float(total2.reshape(-1)[0].real), float(total2.reshape(-1)[0].imag), comm.nccl_version()]))
multi.close(); comm.close()
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
    line2 = [l for l in p.stdout.splitlines() if l.startswith("RESULT2")][0]
    re2, im2, ver = json.loads(line2[len("RESULT2 "):])
    assert abs(complex(re2, im2) - want) / abs(want) < 1e-5 and ver >= 20000


def test_single_process_two_devices(data_dir):
    """jb_multi over two GPUs of ONE process (what Jet::SlicedContractor(devices=...) and the task-based
    contractor with JET_B200_DEVICES use): same sum as one device; also exercises the per-device opt-in of
    >48 KB dynamic shared memory on the second device (tensor-core GEMMs inside ops.gemm)."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from jet_b200 import MultiPlan, NetworkFile, ops
    from jet_b200._lib import check, lib
    net = NetworkFile.load(os.path.join(data_dir, "m10.json"), np.complex64)
    sliced = "p7 s7 h4 m1 m2 I2".split()
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "amplitudes.json")))["m10_s6_sum64_complex128"]
    want = complex(gold["re"], gold["im"])
    with MultiPlan(net, sliced, lanes=2, devices=[0, 1]) as two, MultiPlan(net, sliced, lanes=1, devices=[0]) as one:
        a = two.amplitude().reshape(-1)[0]
        b = one.amplitude().reshape(-1)[0]
        assert abs(a - want) / abs(want) < 1e-5 and abs(a - b) / abs(b) < 1e-10
        assert two.amplitude().reshape(-1)[0] == a
    rng = np.random.default_rng(3)
    ga = (rng.standard_normal((512, 256)) + 1j * rng.standard_normal((512, 256))).astype(np.complex64)
    gb = (rng.standard_normal((256, 256)) + 1j * rng.standard_normal((256, 256))).astype(np.complex64)
    want_g = ga.astype(np.complex128) @ gb.astype(np.complex128)
    for dev in (0, 1, 0):
        check(lib().jb_set_device(dev))
        for x, y, tol in ((ga, gb, 1e-5), (ga.astype(np.complex128), gb.astype(np.complex128), 1e-12)):
            got = ops.gemm(x, y)
            assert np.linalg.norm(got - want_g) / np.linalg.norm(want_g) < tol, dev
    check(lib().jb_set_device(0))
