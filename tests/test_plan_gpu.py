"""GPU parity tests for the contraction plan (whole sliced network on the device) against the
oracle and the golden amplitudes produced by the unmodified reference."""
import json
import os

import numpy as np
import pytest

from oracle import jet_oracle as jo

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
M10_SLICED = "p7 s7 h4 m1 m2 I2 V4 z2 t4 C1".split()
M12_SLICED = "h5 m H10 w y J S G10 P0".split()


def gold():
    return json.load(open(os.path.join(GOLDEN, "amplitudes.json")))


def rel(x, ref):
    return abs(complex(x) - complex(ref)) / abs(complex(ref))


def small_network(dtype, rng, dims=2):
    """A closed 8-tensor network with a fixed path (hand-made; exercises all plan features)."""
    real = np.float32 if dtype == np.complex64 else np.float64
    edges = {"a": (0, 1), "b": (1, 2), "c": (2, 3), "d": (3, 0), "e": (0, 4), "f": (1, 5), "g": (2, 6), "h": (3, 7),
             "i": (4, 5), "j": (5, 6), "k": (6, 7), "l": (7, 4)}
    tensors = []
    for t in range(8):
        idx = [e for e, (u, v) in edges.items() if t in (u, v)]
        n = dims ** len(idx)
        arr = (rng.uniform(-1, 1, n).astype(real) + 1j * rng.uniform(-1, 1, n).astype(real)).astype(dtype)
        tensors.append((idx, arr.reshape([dims] * len(idx))))
    path = [(0, 1), (2, 3), (4, 5), (6, 7), (8, 9), (10, 11), (12, 13)]
    return tensors, path


@pytest.mark.parametrize("dtype", [np.complex64, np.complex128])
@pytest.mark.parametrize("dims", [2, 4, 3])
def test_small_network_sliced_vs_oracle(dtype, dims):
    from jet_b200 import ContractionPlan, NetworkFile
    rng = np.random.default_rng(dims)
    tensors, path = small_network(dtype, rng, dims)
    onet = jo.Network(tensors, path)
    tol = 1e-5 if dtype == np.complex64 else 1e-12
    for sliced in ([], ["a"], ["c", "j"], ["l", "a", "g"]):
        ref = jo.amplitude(onet, sliced).reshape(-1)[0]
        for graph in (True, False):
            with ContractionPlan(NetworkFile(tensors, path), sliced, use_graph=graph, store_results=True) as plan:
                assert plan.num_slices == dims ** len(sliced)
                got = plan.amplitude().reshape(-1)[0]
                assert rel(got, ref) < tol, (sliced, graph)
                # per-slice results in slice order
                for v in range(plan.num_slices):
                    r = jo.amplitude(onet, sliced, [v]).reshape(-1)[0]
                    assert rel(plan.slice_result(v).reshape(-1)[0], r) < 10 * tol
                # explicit slice list, reversed order
                if sliced:
                    ids = list(range(plan.num_slices))[::-1]
                    assert rel(plan.amplitude(ids).reshape(-1)[0], ref) < tol


def test_open_network_result_tensor_and_intermediates():
    """Open output indices: the result is a tensor, reduced elementwise over slices
    (test/Test_TaskBasedContractor.cpp:473-491); KEEP_INTERMEDIATES exposes every step."""
    from jet_b200 import ContractionPlan, NetworkFile
    rng = np.random.default_rng(4)
    tensors, path = small_network(np.complex64, rng, 2)
    tensors[0] = (tensors[0][0] + ["out0"], np.stack([tensors[0][1], 2 * tensors[0][1]], axis=-1))
    tensors[6] = (["out1"] + tensors[6][0], np.stack([tensors[6][1], -tensors[6][1], 1j * tensors[6][1]], axis=0))
    onet = jo.Network(tensors, path)
    nodes = onet.contract(keep_steps=True)
    with ContractionPlan(NetworkFile(tensors, path), [], keep_intermediates=True) as plan:
        got = plan.amplitude()
        assert plan.result_indices == nodes[-1][0]
        assert np.linalg.norm(got - nodes[-1][1]) / np.linalg.norm(nodes[-1][1]) < 1e-5
        for n in range(len(tensors), len(nodes)):
            x = plan.node(n)
            ref = nodes[n][1].reshape(-1)
            assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-5, n
    sliced = ["b", "k"]
    ref = jo.amplitude(onet, sliced)
    with ContractionPlan(NetworkFile(tensors, path), sliced) as plan:
        got = plan.amplitude()
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-5


def test_plan_errors():
    from jet_b200 import ContractionPlan, JetB200Error, NetworkFile
    rng = np.random.default_rng(1)
    tensors, path = small_network(np.complex64, rng)
    with pytest.raises(ValueError, match="Sliced index does not exist."):
        ContractionPlan(NetworkFile(tensors, path), ["nope"])
    with pytest.raises(JetB200Error, match="Node ID 2 in contraction pair is invalid."):
        ContractionPlan(NetworkFile(tensors, [(0, 99)]))


@pytest.mark.parametrize("dt", ["complex64", "complex128"])
def test_m10_sliced_matches_reference_goldens(data_dir, dt):
    from jet_b200 import ContractionPlan, NetworkFile
    g = gold()
    net = NetworkFile.load(os.path.join(data_dir, "m10.json"), np.dtype(dt))
    tol = 1e-5 if dt == "complex64" else 1e-12
    with ContractionPlan(net, M10_SLICED[:6], store_results=True) as plan:
        assert plan.num_slices == 64
        assert plan.stats.jet_flops_per_slice == g[f"m10_s6_slice0_{dt}"]["jet_flops"]
        total = plan.amplitude().reshape(-1)[0]
        for v in range(4):
            e = g[f"m10_s6_slice{v}_{dt}"]
            assert rel(plan.slice_result(v).reshape(-1)[0], complex(e["re"], e["im"])) < tol, v
        e = g[f"m10_s6_sum64_{dt}"]
        assert rel(total, complex(e["re"], e["im"])) < tol
        # complex64 result also within tolerance of the complex128 reference
        e = g["m10_s6_sum64_complex128"]
        assert rel(total, complex(e["re"], e["im"])) < 1e-5
    with ContractionPlan(net, M10_SLICED[:10]) as plan:
        assert plan.num_slices == 1024
        plan.reset()
        plan.run(0, 16)
        e = g[f"m10_s10_sum_first16_{dt}"]
        assert rel(plan.result().reshape(-1)[0], complex(e["re"], e["im"])) < tol


def test_m10_unsliced_full_amplitude(data_dir):
    from jet_b200 import ContractionPlan, NetworkFile
    g = gold()
    net = NetworkFile.load(os.path.join(data_dir, "m10.json"), np.complex64)
    with ContractionPlan(net, []) as plan:
        got = plan.amplitude().reshape(-1)[0]
    e = g["m10_s6_sum64_complex128"]  # the full amplitude equals the sum over all slices
    assert rel(got, complex(e["re"], e["im"])) < 1e-5
    if "m10_full_complex64" in g:
        e = g["m10_full_complex64"]
        assert rel(got, complex(e["re"], e["im"])) < 1e-5


@pytest.mark.parametrize("tot", [0, 10, 30, 60])
def test_gbs_fock4_c128_matches_reference_goldens(data_dir, tot):
    from jet_b200 import ContractionPlan, NetworkFile
    g = gold()
    fn = os.path.join(data_dir, f"gbs_dim2_nc1_lw8_rp5_fock4_total{tot}_0.kraken.json")
    for dt, tol in (("complex128", 1e-12), ("complex64", 1e-5)):
        net = NetworkFile.load(fn, np.dtype(dt))
        with ContractionPlan(net, []) as plan:
            got = plan.amplitude().reshape(-1)[0]
        e = g[f"gbs_fock4_total{tot}_{dt}"]
        assert rel(got, complex(e["re"], e["im"])) < tol, (tot, dt)
    # sliced over two dim-4 indices: 16 slices must sum to the same amplitude
    net = NetworkFile.load(fn, np.complex128)
    dims = net.index_dims()
    sliced = sorted(dims)[:2]
    with ContractionPlan(net, sliced) as plan:
        assert plan.num_slices == 16
        got = plan.amplitude().reshape(-1)[0]
    e = g[f"gbs_fock4_total{tot}_complex128"]
    assert rel(got, complex(e["re"], e["im"])) < 1e-11


def test_m12_single_slices_match_reference_goldens(data_dir):
    from jet_b200 import ContractionPlan, NetworkFile
    g = gold()
    if "m12_s9_slice0_complex64" not in g:
        pytest.skip("heavy goldens not generated")
    net = NetworkFile.load(os.path.join(data_dir, "m12.json"), np.complex64)
    with ContractionPlan(net, M12_SLICED, store_results=True) as plan:
        assert plan.num_slices == 512
        plan.reset()
        plan.run(0, 2)
        for v in (0, 1):
            e = g[f"m12_s9_slice{v}_complex64"]
            assert rel(plan.slice_result(v).reshape(-1)[0], complex(e["re"], e["im"])) < 1e-5, v


def test_lane_plans_match_single_plan(data_dir):
    """LanePlans (several slices in flight on one GPU, one arena + stream + graph per lane) must give
    the same sum as one plan running the same slice ids (FP64 accumulators; the only difference is
    the order in which the per-slice results are added)."""
    from jet_b200 import ContractionPlan, LanePlans, NetworkFile
    net = NetworkFile.load(os.path.join(data_dir, "m10.json"), np.complex64)
    sliced = ["p7", "s7", "h4", "m1", "m2", "I2"]
    ids = list(range(0, 64, 3))
    with ContractionPlan(net, sliced) as one, LanePlans(net, sliced, lanes=3) as lanes:
        a = one.amplitude(ids).reshape(-1)[0]
        b = lanes.amplitude(ids).reshape(-1)[0]
        assert abs(a - b) / abs(a) < 1e-12
        b2 = lanes.amplitude(ids).reshape(-1)[0]
        assert b2 == b  # deterministic for a fixed lane count


def test_plans_beyond_the_constant_bank_slots_stay_fused(data_dir):
    """A device has five constant-bank slots for the matrices of the chains' register stages.  The sixth and later
    live plans must still fuse their chains (matrices in shared memory, no register stage) and give the same sums
    within complex64 accuracy — not fall back to one kernel per step."""
    from jet_b200 import ContractionPlan, NetworkFile
    net = NetworkFile.load(os.path.join(data_dir, "m10.json"), np.complex64)
    sliced = ["p7", "s7", "h4", "m1", "m2", "I2"]
    g = gold()
    e = g["m10_s6_slice0_complex128"]
    want = complex(e["re"], e["im"])
    plans = [ContractionPlan(net, sliced) for _ in range(7)]
    try:
        chains = [int(p.stats.chains) for p in plans]
        assert min(chains) > 0 and chains[-1] == chains[0], chains
        regs = [sum(u.register_steps for u in p.ops()) for p in plans]
        assert regs[0] >= regs[-1] and regs[-1] == 0, regs  # the last plans found no slot: no register stage
        for p in (plans[0], plans[-1]):
            assert rel(p.amplitude([0]).reshape(-1)[0], want) < 1e-5
    finally:
        for p in plans:
            p.close()


def test_fully_sliced_leaf_scalar_chain():
    """Regression (found by tools/stress_plan_gpu.py): when every index of a leaf is sliced, the chained
    tensor of a fused run is a scalar and the chain's shared-memory tile holds ONE element; the step matrices
    behind it must stay 16-byte aligned."""
    from jet_b200 import ContractionPlan, LanePlans, NetworkFile
    rng = np.random.default_rng(174)

    def rc(shape):
        n = int(np.prod(shape))
        return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex64).reshape(shape)

    tensors = [(["e0", "e3", "e4"], rc([2, 2, 2])), (["e1", "e0"], rc([2, 2])), (["e1", "e4", "e3", "e2"], rc([2] * 4)),
               (["e2"], rc([2]))]
    path = [(3, 2), (0, 4), (5, 1)]
    sliced = ["e3", "e1", "e0", "e4"]
    ref = np.asarray(jo.amplitude(jo.Network(tensors, path), sliced)).reshape(-1)[0]
    for make in (lambda: ContractionPlan(NetworkFile(tensors, path), sliced),
                 lambda: LanePlans(NetworkFile(tensors, path), sliced, lanes=2)):
        with make() as plan:
            got = plan.amplitude().reshape(-1)[0]
            assert rel(got, ref) < 1e-5


def test_gbs_fock8_c128_matches_reference_golden(data_dir):
    """The compute-bound complex128 workload (dim-8 indices, 2^27-element intermediates: DMMA GEMMs with gathered
    A, split-K, deferred slicing) against the reference's complex128 amplitude of the same file
    (tools/make_fock8_golden.py), all 64 slices of the two greedily chosen indices."""
    gpath = os.path.join(os.path.dirname(__file__), "golden", "amplitudes_fock8.json")
    if not os.path.exists(gpath):
        pytest.skip("fock8 golden not generated")
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from jet_b200 import ContractionPlan
    g = json.load(open(gpath))["gbs_fock8_total0_complex128"]
    want = complex(g["re"], g["im"])
    net, sliced, dt, _ = bench.load_network("gbs_fock8_total0_s2")
    assert dt == "complex128"
    with ContractionPlan(net, sliced) as plan:
        assert plan.num_slices == 64
        got = complex(plan.amplitude().reshape(-1)[0])
    assert abs(got - want) / abs(want) < 1e-12, (got, want)


def test_slice_batching_is_bit_identical(data_dir):
    """Slice batching (several slices per launch: per-slice tensors replicated `batch` times in the arena, kernels
    take the slice from blockIdx.y/z) must not change a single bit: same arithmetic per slice, same accumulation
    order.  m10 / s=10 (batch 64 by default), uneven counts (whole batches + remainder), explicit id lists, per-slice
    results, and a random network with non-power-of-two extents (TTGT units are looped, not batched)."""
    from jet_b200 import ContractionPlan, NetworkFile
    net = NetworkFile.load(os.path.join(data_dir, "m10.json"), np.complex64)
    sliced = "p7 s7 h4 m1 m2 I2 V4 z2 t4 C1".split()
    with ContractionPlan(net, sliced, batch=1, store_results=True) as one, \
            ContractionPlan(net, sliced, store_results=True) as auto, \
            ContractionPlan(net, sliced, batch=8, store_results=True) as eight:
        assert one.stats.batch == 1 and auto.stats.batch == 64 and eight.stats.batch == 8
        for first, count in ((0, 200), (37, 75), (1000, 24), (5, 3)):
            want = None
            for plan in (one, auto, eight):
                plan.reset()
                plan.run(first, count)
                got = plan.result().reshape(-1)[0]
                per = [complex(plan.slice_result(k).reshape(-1)[0]) for k in (0, count // 2, count - 1)]
                if want is None:
                    want = (got, per)
                else:
                    assert got == want[0] and per == want[1], (first, count, plan.stats.batch)
        ids = [3, 1000, 17, 17, 512, 9, 77, 640, 2, 1023, 0]
        ref = one.amplitude(ids).reshape(-1)[0]
        assert auto.amplitude(ids).reshape(-1)[0] == ref and eight.amplitude(ids).reshape(-1)[0] == ref
    rng = np.random.default_rng(11)

    def rc(shape):
        n = int(np.prod(shape))
        return (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(np.complex128).reshape(shape)

    tensors = [(["a", "b", "s"], rc([3, 4, 2])), (["b", "c", "t"], rc([4, 2, 4])), (["c", "d", "s"], rc([2, 5, 2])),
               (["d", "a", "t", "u"], rc([5, 3, 4, 2])), (["u", "o"], rc([2, 3]))]
    path = [(0, 1), (2, 3), (5, 6), (7, 4)]
    with ContractionPlan(NetworkFile(tensors, path), ["s", "t"], batch=1) as one, \
            ContractionPlan(NetworkFile(tensors, path), ["s", "t"], batch=4) as four:
        assert four.stats.batch == 4
        a, b = one.amplitude(), four.amplitude()
        assert np.array_equal(a, b)
        want = np.asarray(jo.amplitude(jo.Network(tensors, path), ["s", "t"]))
        assert np.linalg.norm(b.reshape(-1) - want.reshape(-1)) / np.linalg.norm(want) < 1e-12
