"""The Python circuit front end (jet_b200/{gate,state,circuit}.py) against the reference's.

CPU part: every qubit gate matrix element by element against the reference's own ``_data()`` output
(tests/golden/gates.npz, written by tools/make_gate_golden.py from /root/reference/python/jet/gate.py), the Fock gates
and the bookkeeping against the known answers of the reference's tests (python/tests/test_gate.py,
test_state.py, test_circuit.py — restated, the expected values are the reference's), and the circuit -> tensor network
lowering against a plain state-vector simulation through the numpy oracle.
GPU part: random circuits contracted by the plan engine (``Circuit.amplitude``) and through the drop-in
``TensorNetwork`` / ``TaskBasedContractor`` bindings against the same state-vector simulation
(1e-12 complex128 / 1e-5 complex64)."""
import os
from math import pi, sqrt

import numpy as np
import pytest

from jet_b200 import circuit as jc
from jet_b200 import gate as jg
from jet_b200 import state as js
from oracle import jet_oracle as jo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gates.npz")
R2 = 1 / sqrt(2)


# ---------------------------------------------------------------- gates
def test_qubit_gate_matrices_match_the_reference():
    z = np.load(GOLD)
    keys = [k for k in z.files if k.endswith("/matrix")]
    assert len(keys) == 53
    for key in keys:
        cls, rep, _ = key.split("/")
        gate = getattr(jg, cls)(*z[f"{cls}/{rep}/params"])
        assert gate.num_wires == int(z[f"{cls}/{rep}/num_wires"])
        assert gate.name == cls and gate.dimension == 2
        assert np.abs(gate._data() - z[key]).max() < 1e-14, key
        assert np.abs(jg.Adjoint(gate)._data() - z[f"{cls}/{rep}/adjoint"]).max() < 1e-14, key
        assert list(gate.params or []) == pytest.approx(list(z[f"{cls}/{rep}/params"]))


def test_gate_registry_has_the_reference_names():
    z = np.load(GOLD)
    want = dict(zip(z["registry/names"].tolist(), z["registry/classes"].tolist()))
    have = {name: cls.__name__ for name, cls in jg.GateFactory.registry.items()}
    assert have == want


@pytest.mark.parametrize("gate, column, want", [
    # python/tests/test_gate.py:258-381 (thewalrus matrices): gate applied to a Fock basis state
    (lambda: jg.Displacement(2, pi / 2, 3), 0, [0.135335283237, 0.270670566473j, -0.382785986042]),
    (lambda: jg.Displacement(2, pi / 2, 3), 1, [0.270670566473j, -0.40600584971, -0.382785986042j]),
    (lambda: jg.Displacement(2, pi / 2, 3), 2, [-0.382785986042, -0.382785986042j, 0.135335283237]),
    (lambda: jg.Squeezing(2, pi / 2, 3), 0, [0.515560111756, 0, -0.351442087775j]),
    (lambda: jg.Squeezing(2, pi / 2, 3), 1, [0, 0.137037026803, 0]),
    (lambda: jg.Squeezing(2, pi / 2, 3), 2, [-0.351442087775j, 0, -0.203142935143]),
    (lambda: jg.TwoModeSqueezing(3, pi / 4, 2), 0, [0.099327927419, 0, 0, 0.069888119434 + 0.069888119434j]),
    (lambda: jg.TwoModeSqueezing(3, pi / 4, 2), 1, [0, 0.009866037165, 0, 0]),
    (lambda: jg.TwoModeSqueezing(3, pi / 4, 2), 2, [0, 0, 0.009866037165, 0]),
    (lambda: jg.TwoModeSqueezing(3, pi / 4, 2), 3, [-0.069888119434 + 0.069888119434j, 0, 0, -0.097367981372]),
    (lambda: jg.Beamsplitter(pi / 4, pi / 2, 2), 0, [1, 0, 0, 0]),
    (lambda: jg.Beamsplitter(pi / 4, pi / 2, 2), 1, [0, R2, R2 * 1j, 0]),
    (lambda: jg.Beamsplitter(pi / 4, pi / 2, 2), 2, [0, R2 * 1j, R2, 0]),
    (lambda: jg.Beamsplitter(pi / 4, pi / 2, 2), 3, [0, 0, 0, 0]),
])
def test_fock_gates_known_answers(gate, column, want):
    assert gate()._data()[:, column] == pytest.approx(np.asarray(want, dtype=np.complex128), abs=1e-11)


def test_fock_gates_are_exact_inside_the_cutoff():
    """A gate truncated at cutoff c must equal the top-left block of the same gate at a larger cutoff (the disentangled
    forms lose nothing inside the cutoff), and the untruncated operators are unitary: at a generous cutoff the low block
    of U^dag U is the identity."""
    for make in (lambda c: jg.Displacement(0.4, 0.7, c), lambda c: jg.Squeezing(0.3, -1.1, c)):
        small, big = make(4)._data(), make(24)._data()
        assert np.abs(small - big[:4, :4]).max() < 1e-13
        assert np.abs((big.conj().T @ big)[:4, :4] - np.eye(4)).max() < 1e-8
    for make in (lambda c: jg.TwoModeSqueezing(0.3, 0.5, c), lambda c: jg.Beamsplitter(0.6, -0.4, c)):
        small, big = make(3)._data().reshape(3, 3, 3, 3), make(12)._data().reshape(12, 12, 12, 12)
        assert np.abs(small - big[:3, :3, :3, :3]).max() < 1e-13
    bs = jg.Beamsplitter(0.6, -0.4, 5)._data()
    # photon-number conserving: inputs with at most 4 photons in all stay inside the cutoff, so those columns are orthonormal
    cols = [k * 5 + l for k in range(5) for l in range(5) if k + l <= 4]
    assert np.abs((bs.conj().T @ bs)[np.ix_(cols, cols)] - np.eye(len(cols))).max() < 1e-12


def test_gate_validation_and_factory():
    with pytest.raises(ValueError, match="The dimension of a qubit gate must be exactly two."):
        jg.Hadamard().dimension = 3
    with pytest.raises(ValueError, match="The dimension of a Fock gate must be greater than one."):
        jg.Displacement(1, 2, cutoff=1)
    gate = jg.CX()
    assert gate.indices is None
    gate.indices = ["a", "b", "c", "d"]
    assert gate.indices == ["a", "b", "c", "d"]
    gate.indices = None
    for bad in (1, ["a", "b", "c", 4], ["a", "a", "b", "c"], "abcd"):
        with pytest.raises(ValueError, match="Indices must be a sequence of unique strings."):
            gate.indices = bad
    with pytest.raises(ValueError, match="Gates must have two indices per wire; received 3 indices for 2 wires."):
        gate.indices = ["a", "b", "c"]
    # factory (python/tests/test_gate.py: GateFactory tests)
    with pytest.raises(KeyError, match="The name 'nope' does not exist in the gate registry."):
        jg.GateFactory.create("nope")
    rx = jg.GateFactory.create("rx", 0.3, adjoint=True, scalar=2)
    assert np.abs(rx._data() - 2 * jg.RX(0.3)._data().conj().T).max() < 1e-15
    assert rx.name == "RX" and rx.params == [0.3] and rx.num_wires == 1
    assert isinstance(jg.GateFactory.create("D", 1, 2, cutoff=3), jg.Displacement)
    with pytest.raises(KeyError, match="already exist in the gate registry"):
        jg.GateFactory.register(names=["X"])(type("Mine", (jg.QubitGate,), {"_data": lambda self: None}))
    with pytest.raises(ValueError, match="is not a subclass of Gate"):
        jg.GateFactory.register(names=["fresh"])(int)

    @jg.GateFactory.register(names=["MyGate", "mygate"])
    class MyGate(jg.QubitGate):
        def __init__(self):
            super().__init__(name="MyGate", num_wires=1)

        def _data(self):
            return np.eye(2)

    assert isinstance(jg.GateFactory.create("mygate"), MyGate)
    jg.GateFactory.unregister(MyGate)
    assert "MyGate" not in jg.GateFactory.registry and "mygate" not in jg.GateFactory.registry


def test_gate_and_state_tensors():
    """tensor(): default labels "0", "1", ...; shape [dim] * rank; row-major matrix data (python/tests/test_gate.py,
    test_state.py).  Uses the compiled bindings (no GPU work)."""
    t = jg.CX().tensor(dtype=np.complex64)
    assert t.indices == ["0", "1", "2", "3"] and t.shape == [2, 2, 2, 2]
    assert np.asarray(t.data) == pytest.approx(jg.CX()._data().reshape(-1))
    g = jg.Beamsplitter(0.1, 0.2, 3)
    g.indices = ["a", "b", "c", "d"]
    t = g.tensor()
    assert t.indices == ["a", "b", "c", "d"] and t.shape == [3, 3, 3, 3]
    s = js.QubitRegister(2, data=np.array([0, 1, 0, 0]))
    s.indices = ["x", "y"]
    t = s.tensor()
    assert t.indices == ["x", "y"] and t.shape == [2, 2] and np.asarray(t.data) == pytest.approx([0, 1, 0, 0])
    assert js.Qudit(dim=3).tensor().shape == [3]


# ---------------------------------------------------------------- states
def test_states():
    assert js.Qubit().name == "Qubit" and js.Qudit(3).name == "Qudit(d=3)"
    assert js.QubitRegister(2).name == "Qubit[2]" and js.QuditRegister(3, 2).name == "Qudit(d=3)[2]"
    assert js.Qudit(3)._data() == pytest.approx([1, 0, 0])
    assert js.QuditRegister(3, 2)._data() == pytest.approx([1] + [0] * 8)
    assert js.Qubit(data=np.array([[0], [1]]))._data() == pytest.approx([0, 1])
    assert js.Qubit() == js.Qudit(2) and js.Qubit() != js.Qubit(data=np.array([0, 1]))
    s = js.QubitRegister(2)
    assert s.num_wires == 2 and s.indices is None
    s.indices = ["a", "b"]
    with pytest.raises(ValueError, match="Indices must be a sequence of unique strings."):
        s.indices = ["a", "a"]
    with pytest.raises(ValueError, match="States must have one index per wire. Received 3 indices for 2 wires."):
        s.indices = ["a", "b", "c"]


# ---------------------------------------------------------------- circuits: bookkeeping (python/tests/test_circuit.py)
def test_circuit_bookkeeping_and_errors():
    assert jc.Wire(1, depth=2).index == "1-2"
    c = jc.Circuit(num_wires=4)
    assert c.dimension == 2
    assert [w.id_ for w in c.wires] == [0, 1, 2, 3]
    ops = list(c.operations)
    assert [op.wire_ids for op in ops] == [[0], [1], [2], [3]]
    assert all(isinstance(op.part, js.Qudit) and op.part.indices == [f"{i}-0"] for i, op in enumerate(ops))
    with pytest.raises(ValueError, match=r"Wire ID 4 falls outside the range \[0, 4\)."):
        c.append_gate(jg.Hadamard(), [4])
    with pytest.raises(ValueError, match="Wire ID 0 is specified more than once."):
        c.append_gate(jg.CX(), [0, 0])
    with pytest.raises(ValueError, match=r"Number of wire IDs \(1\) must match the number of wires connected to the gate \(2\)."):
        c.append_gate(jg.CX(), [0])
    with pytest.raises(ValueError, match=r"Number of wire IDs \(2\) must match the number of wires connected to the state \(1\)."):
        c.append_state(js.Qubit(), [0, 1])
    h = jg.Hadamard()
    c.append_gate(h, [3])
    assert h.indices == ["3-1", "3-0"] and list(c.indices([3])) == ["3-1"]
    cz = jg.CZ()
    c.append_gate(cz, [3, 1])
    assert cz.indices == ["3-2", "1-1", "3-1", "1-0"]
    assert list(c.indices([1, 2, 3])) == ["1-1", "2-0", "3-2"]
    reg = js.QubitRegister(2)
    c.append_state(reg, [1, 3])
    assert reg.indices == ["1-1", "3-2"]
    assert [w.closed for w in c.wires] == [False, True, False, True]
    with pytest.raises(ValueError, match="Wire 1 is closed."):
        c.append_gate(jg.PauliX(), [1])
    assert [type(op.part).__name__ for op in c.operations][4:] == ["Hadamard", "CZ", "QuditRegister"]


def _statevector(circuit):
    """Plain simulation of the circuit's gates on |0...0> (test-side oracle): amplitudes indexed by wire, wire 0 slowest."""
    n = sum(1 for _ in circuit.wires)
    d = circuit.dimension
    psi = np.zeros([d] * n, dtype=np.complex128)
    psi[(0,) * n] = 1
    for op in list(circuit.operations)[n:]:
        if isinstance(op.part, jg.Gate):
            k = len(op.wire_ids)
            m = np.asarray(op.part._data(), dtype=np.complex128).reshape([d] * (2 * k))
            psi = np.tensordot(m, psi, axes=(list(range(k, 2 * k)), list(op.wire_ids)))
            psi = np.moveaxis(psi, list(range(k)), list(op.wire_ids))
    return psi


def _contract_with_oracle(circuit, dtype=np.complex128):
    net = circuit.network_file(dtype)
    return jo.amplitude(jo.Network(net.tensors, net.path), [])


def _random_circuit(rng, n, depth):
    c = jc.Circuit(num_wires=n)
    one = [lambda: jg.Hadamard(), lambda: jg.SX(), lambda: jg.T(), lambda: jg.RX(rng.uniform(-3, 3)),
           lambda: jg.U3(*rng.uniform(-3, 3, 3)), lambda: jg.Rot(*rng.uniform(-3, 3, 3)), lambda: jg.PauliY()]
    two = [lambda: jg.CX(), lambda: jg.CZ(), lambda: jg.ISWAP(), lambda: jg.CRY(rng.uniform(-3, 3)),
           lambda: jg.CPhaseShift(rng.uniform(-3, 3)), lambda: jg.SWAP()]
    for _ in range(depth):
        for w in range(n):
            c.append_gate(one[rng.integers(len(one))](), [w])
        order = rng.permutation(n)
        for a, b in zip(order[0::2], order[1::2]):
            c.append_gate(two[rng.integers(len(two))](), [int(a), int(b)])
    if n >= 3:
        c.append_gate(jg.Toffoli(), [int(i) for i in rng.permutation(n)[:3]])
    return c


def test_circuit_lowering_matches_statevector_on_cpu():
    """leaves + searched path contracted by the numpy oracle == plain state-vector simulation: an amplitude <b|U|0>
    (every wire closed with a basis state) and an open circuit (the whole output state, up to index order)."""
    rng = np.random.default_rng(7)
    c = _random_circuit(rng, 5, 3)
    psi = _statevector(c)
    bits = [1, 0, 1, 1, 0]
    for w, b in enumerate(bits):
        c.append_state(js.Qubit(data=np.eye(2)[b]), [w])
    got = np.asarray(_contract_with_oracle(c)).reshape(-1)[0]
    assert abs(got - psi[tuple(bits)]) < 1e-12
    # open output: labels of the result tell the wire order
    c2 = _random_circuit(np.random.default_rng(8), 4, 2)
    net = c2.network_file()
    labels, tensor = jo.Network(net.tensors, net.path).contract()
    wires = [int(lbl.split("-")[0]) for lbl in labels]
    assert sorted(wires) == [0, 1, 2, 3]
    assert np.abs(np.transpose(tensor, np.argsort(wires)) - _statevector(c2)).max() < 1e-12


@pytest.mark.parametrize("operations, observable, want", [
    # python/tests/test_circuit.py:120-167
    ([], [("Z", [0])], 1),
    ([("X", [0])], [("Z", [0])], -1),
    ([("RX", [0], 1), ("RY", [1], 2), ("CNOT", [0, 1])], [("Y", [0]), ("X", [1])], -0.8414709848078962),
])
def test_take_expected_value_on_cpu(operations, observable, want):
    c = jc.Circuit(num_wires=2)
    for name, wires, *params in operations:
        c.append_gate(jg.GateFactory.create(name, *params), wires)
    n_before = len(list(c.operations))
    c.take_expected_value([jc.Operation(part=jg.GateFactory.create(name), wire_ids=wires) for name, wires in observable])
    ops = list(c.operations)
    assert len(ops) == n_before + len(observable) + len(operations) + 2
    assert all(w.closed for w in c.wires)
    got = np.asarray(_contract_with_oracle(c)).reshape(-1)[0]
    assert got.real == pytest.approx(want) and abs(got.imag) < 1e-12


# ---------------------------------------------------------------- GPU: the engine contracts what the front end builds
@pytest.mark.gpu
@pytest.mark.parametrize("dtype, tol", [(np.complex128, 1e-12), (np.complex64, 1e-5)])
def test_circuit_amplitudes_on_the_plan_engine(dtype, tol):
    rng = np.random.default_rng(11)
    for n, depth in [(4, 2), (8, 4), (12, 5)]:
        c = _random_circuit(rng, n, depth)
        psi = _statevector(c)
        bits = [int(b) for b in rng.integers(0, 2, n)]
        for w, b in enumerate(bits):
            c.append_state(js.Qubit(data=np.eye(2)[b]), [w])
        want = psi[tuple(bits)]
        got = c.amplitude(dtype=dtype)
        scale = max(abs(want), 2.0 ** (-n / 2))  # amplitudes of a random circuit are ~2^(-n/2): compare on that scale
        assert abs(got - want) / scale < tol * 10, (n, got, want)
        # sliced over two interior indices: the FP64 sum over the slices is the same amplitude
        net = c.network_file(dtype)
        counts = {}
        for idx, _ in net.tensors:
            for i in idx:
                counts[i] = counts.get(i, 0) + 1
        interior = sorted(i for i, k in counts.items() if k == 2)[n:n + 2]
        got_sliced = c.amplitude(dtype=dtype, sliced=interior)
        assert abs(got_sliced - want) / scale < tol * 10


@pytest.mark.gpu
def test_open_circuit_state_on_the_plan_engine():
    """A circuit with open wires: the engine returns the output state as (labels, tensor)."""
    c = _random_circuit(np.random.default_rng(21), 6, 3)
    labels, tensor = c.amplitude(dtype=np.complex128)
    wires = [int(lbl.split("-")[0]) for lbl in labels]
    assert sorted(wires) == list(range(6))
    assert np.abs(np.transpose(tensor, np.argsort(wires)) - _statevector(c)).max() < 1e-12


@pytest.mark.gpu
def test_circuit_through_the_dropin_bindings():
    """The reference's own usage: circuit.tensor_network() -> contract / TaskBasedContractor
    (python/tests/test_circuit.py:189-240)."""
    from jet_b200 import jet

    c = jet.Circuit(num_wires=1)
    c.append_gate(jet.PauliX(), wire_ids=[0])
    t = c.tensor_network().contract()
    assert t.indices == ["0-1"] and t.shape == [2] and np.asarray(t.data) == pytest.approx([0, 1])
    c = jet.Circuit(num_wires=2)
    c.append_gate(jet.GateFactory.create("H"), wire_ids=[0])
    c.append_gate(jet.GateFactory.create("CNOT"), wire_ids=[0, 1])
    t = c.tensor_network().contract()
    assert t.indices == ["0-2", "1-1"] and t.shape == [2, 2]
    assert np.asarray(t.data) == pytest.approx([R2, 0, 0, R2])
    # expected value through the task-based contractor
    c = jet.Circuit(num_wires=2)
    c.append_gate(jet.GateFactory.create("RX", 1), [0])
    c.append_gate(jet.GateFactory.create("RY", 2), [1])
    c.append_gate(jet.GateFactory.create("CNOT"), [0, 1])
    c.take_expected_value([jet.Operation(part=jet.GateFactory.create("Y"), wire_ids=[0]),
                           jet.Operation(part=jet.GateFactory.create("X"), wire_ids=[1])])
    tn = c.tensor_network()
    net = c.network_file()
    tbc = jet.TaskBasedContractor()
    tbc.add_contraction_tasks(tn, jet.PathInfo(tn, net.path))
    tbc.contract()
    assert complex(tbc.results[0].scalar).real == pytest.approx(-0.8414709848078962)


@pytest.mark.gpu
def test_simulate_helpers_known_answers_and_statevector():
    """The quantities the reference's interpreter returns (python/tests/test_interpreter.py:309-520, the programs built
    directly as circuits): amplitudes, probabilities, expected values."""
    from jet_b200 import simulate as sim

    # X | [0]; amplitude(state: [0]) -> 0, amplitude(state: [1]) -> 1
    c = jc.Circuit(1)
    c.append_gate(jg.PauliX(), [0])
    assert [sim.compute_amplitude(c, [b]) for b in (0, 1)] == pytest.approx([0, 1])
    assert sim.compute_probabilities(c) == pytest.approx([0, 1])
    assert sim.compute_probabilities(jc.Circuit(1)) == pytest.approx([1, 0])          # one tensor: nothing to contract
    assert sim.compute_probabilities(jc.Circuit(3)) == pytest.approx([1] + [0] * 7)
    # Bell pair
    c = jc.Circuit(2)
    c.append_gate(jg.GateFactory.create("H"), [0])
    c.append_gate(jg.GateFactory.create("CNOT"), [0, 1])
    amps = [sim.compute_amplitude(c, [a, b]) for a in (0, 1) for b in (0, 1)]
    assert amps == pytest.approx([R2, 0, 0, R2])
    assert sim.compute_probabilities(c) == pytest.approx([0.5, 0, 0, 0.5])
    assert len(list(c.operations)) == 4  # the circuit handed in is untouched
    # TwoModeSqueezing(3, 1) | [0, 1] at the default dimension 2 (test_interpreter.py:342-355)
    c = jc.Circuit(2)
    c.append_gate(jg.TwoModeSqueezing(3, 1), [0, 1])
    amps = [sim.compute_amplitude(c, [a, b]) for a in (0, 1) for b in (0, 1)]
    assert amps == pytest.approx([0.0993279274194332, 0, 0, 0.053401711152745175 + 0.08316823745907517j])
    # expected value (test_circuit.py:155-165)
    c = jc.Circuit(2)
    c.append_gate(jg.GateFactory.create("RX", 1), [0])
    c.append_gate(jg.GateFactory.create("RY", 2), [1])
    c.append_gate(jg.GateFactory.create("CNOT"), [0, 1])
    obs = [jc.Operation(part=jg.GateFactory.create("Y"), wire_ids=[0]), jc.Operation(part=jg.GateFactory.create("X"), wire_ids=[1])]
    assert sim.compute_expected_value(c, obs).real == pytest.approx(-0.8414709848078962)
    # random 8-qubit circuit: probabilities and <Z_3> against the state vector, complex64 too
    rc = _random_circuit(np.random.default_rng(31), 8, 4)
    psi = _statevector(rc)
    p = sim.compute_probabilities(rc)
    assert p.real == pytest.approx((np.abs(psi) ** 2).reshape(-1), abs=1e-12) and abs(p.real.sum() - 1) < 1e-12
    z3 = np.sum(np.abs(psi) ** 2 * np.where(np.arange(2).reshape([1, 1, 1, 2, 1, 1, 1, 1]) == 0, 1.0, -1.0))
    got = sim.compute_expected_value(rc, [jc.Operation(part=jg.PauliZ(), wire_ids=[3])])
    assert got.real == pytest.approx(z3, abs=1e-12)
    got32 = sim.compute_expected_value(rc, [jc.Operation(part=jg.PauliZ(), wire_ids=[3])], dtype=np.complex64)
    assert got32.real == pytest.approx(z3, abs=1e-5)
    with pytest.raises(ValueError, match=r"The state has 1 \(!= 8\) entries."):
        sim.compute_amplitude(rc, [0])
