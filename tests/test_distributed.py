"""CPU tests of the N>1 host path: world_size-2 gloo processes partition the slices of the m10
network, each contracts its block (with the numpy oracle standing in for the device here — the
GPU twin of this test is tests/test_distributed_gpu.py) and one reduce gives the amplitude."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from jet_b200.distributed import slice_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, {root!r})
import torch.distributed as dist
from jet_b200.distributed import slice_range, reduce_amplitude
from oracle import jet_oracle as jo
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
net = jo.Network.from_file(os.path.join({root!r}, "data", "_ref", "m10.json"), "complex64")
sliced = "p7 s7 h4 m1 m2 I2".split()
first, count = slice_range(6, world, rank)          # 6 of the 64 slices -> uneven split at world=4
part = jo.amplitude(net, sliced, range(first, first + count))
total = reduce_amplitude(np.asarray(part).reshape(1))
if rank == 0:
    print("RESULT", json.dumps([float(total[0].real), float(total[0].imag), first, count]))
dist.destroy_process_group()
'''


def test_slice_range_partitions_exactly():
    for n in (1, 7, 64, 512, 1000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, count = slice_range(n, world, r)
                seen += list(range(first, first + count))
            assert seen == list(range(n))
            counts = [slice_range(n, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        slice_range(4, 2, 2)


@pytest.mark.parametrize("world", [2])
def test_gloo_world2_reduces_partial_amplitudes(world, tmp_path, data_dir):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")][0]
    re, im, first, count = json.loads(line[len("RESULT "):])
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "amplitudes.json")))
    # slices 0..3 are pinned by the reference goldens; 4 and 5 by the oracle itself
    from oracle import jet_oracle as jo
    net = jo.Network.from_file(os.path.join(data_dir, "m10.json"), "complex64")
    want = jo.amplitude(net, "p7 s7 h4 m1 m2 I2".split(), range(6)).reshape(-1)[0]
    assert abs(complex(re, im) - want) / abs(want) < 1e-6
    assert (first, count) == (0, 3)
    g = sum(complex(gold[f"m10_s6_slice{v}_complex64"]["re"], gold[f"m10_s6_slice{v}_complex64"]["im"]) for v in range(4))
    part4 = jo.amplitude(net, "p7 s7 h4 m1 m2 I2".split(), range(4)).reshape(-1)[0]
    assert abs(part4 - g) / abs(g) < 1e-6
