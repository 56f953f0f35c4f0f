"""Quantum gates as tensors: the gate layer of the reference's Python front end (reference python/jet/gate.py:67-952:
``Gate``, ``GateFactory``, ``Adjoint``, ``Scale``, the Fock gates and the qubit gates), restated over the B200 engine's
``Tensor`` factory (``jet_b200.jet.Tensor``).

Same class names, constructor arguments, registry names and error messages as the reference, so circuits written for
``jet`` build unchanged.  What differs is underneath:

* the matrices are produced by small closed-form builders in this file (a table of constants plus a handful of
  rotation formulas) and checked element by element against the reference's own ``_data()`` output
  (``tests/golden/gates.npz``, written by ``tools/make_gate_golden.py``);
* the four continuous-variable gates do not need ``thewalrus`` (absent here): their Fock matrix elements are computed
  exactly inside the cutoff from the disentangled forms of the operators (below) and pinned by the known answers of
  the reference's tests (python/tests/test_gate.py:258-381).
"""
from __future__ import annotations

import cmath
import math
from abc import ABC, abstractmethod
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

__all__ = [
    "Gate", "GateFactory", "Adjoint", "Scale",
    "FockGate", "Displacement", "Squeezing", "TwoModeSqueezing", "Beamsplitter",
    "QubitGate", "Hadamard", "PauliX", "PauliY", "PauliZ", "S", "T", "SX", "CX", "CY", "CZ", "SWAP", "ISWAP", "CSWAP",
    "Toffoli", "RX", "RY", "RZ", "PhaseShift", "CPhaseShift", "Rot", "CRX", "CRY", "CRZ", "CRot", "U1", "U2", "U3",
]


def _check_indices(indices, expected: int, message: str):
    """Index labels of a gate or state: None, or `expected` unique strings (reference gate.py:129-160)."""
    if indices is None:
        return
    ok = isinstance(indices, Sequence) and not isinstance(indices, str) and all(isinstance(i, str) for i in indices)
    if not ok or len(set(indices)) != len(indices):
        raise ValueError("Indices must be a sequence of unique strings.")
    if len(indices) != expected:
        raise ValueError(message.format(len(indices)))


class Gate(ABC):
    """A gate acting on ``num_wires`` qudits of dimension ``dim`` (reference gate.py:67-191).  Its tensor has one output
    and one input index per wire: shape ``[dim] * 2 * num_wires``, row-major data of the ``dim^n x dim^n`` matrix."""

    def __init__(self, name: str, num_wires: int, dim: int, params: Optional[List[float]] = None):
        self.name = name
        self._indices = None
        self._num_wires = num_wires
        self._params = params
        self._validate_dimension(dim)
        self._dim = dim

    @property
    def dimension(self) -> int:
        return self._dim

    @dimension.setter
    def dimension(self, dim: int) -> None:
        self._validate_dimension(dim)
        self._dim = dim

    @abstractmethod
    def _validate_dimension(self, dim: int) -> None:
        """Raises ValueError when the gate cannot act on qudits of this dimension."""

    @property
    def indices(self) -> Optional[Sequence[str]]:
        return self._indices

    @indices.setter
    def indices(self, indices: Optional[Sequence[str]]) -> None:
        _check_indices(indices, 2 * self._num_wires,
                       "Gates must have two indices per wire; received {} indices for " + f"{self._num_wires} wires.")
        self._indices = indices

    @property
    def num_wires(self) -> int:
        return self._num_wires

    @property
    def params(self) -> Optional[List[float]]:
        return self._params

    @abstractmethod
    def _data(self) -> np.ndarray:
        """The matrix of the gate (rows = outputs)."""

    def tensor(self, dtype: np.dtype = np.complex128):
        """The gate as an engine tensor with labels ``indices`` (default "0", "1", ...: outputs first)."""
        from .jet import Tensor  # the compiled bindings are only needed here

        data = np.asarray(self._data()).reshape(-1)
        indices = list(self.indices) if self.indices is not None else [str(i) for i in range(2 * self._num_wires)]
        dim = int(round(len(data) ** (1.0 / len(indices))))
        return Tensor(indices=indices, shape=[dim] * len(indices), data=data, dtype=dtype)


class GateFactory:
    """Name -> gate class registry (reference gate.py:194-275)."""

    registry: Dict[str, type] = {}

    @staticmethod
    def create(name: str, *params: float, adjoint: bool = False, scalar: float = 1, **kwargs) -> Gate:
        if name not in GateFactory.registry:
            raise KeyError(f"The name '{name}' does not exist in the gate registry.")
        gate = GateFactory.registry[name](*params, **kwargs)
        if adjoint:
            gate = Adjoint(gate=gate)
        if scalar != 1:
            gate = Scale(gate=gate, scalar=scalar)
        return gate

    @staticmethod
    def register(names: Sequence[str]) -> Callable[[type], type]:
        def wrapper(subclass: type) -> type:
            if not (isinstance(subclass, type) and issubclass(subclass, Gate)):
                raise ValueError(f"The type '{subclass.__name__}' is not a subclass of Gate.")
            conflicts = set(names) & set(GateFactory.registry)
            if conflicts:
                raise KeyError(f"The names {conflicts} already exist in the gate registry.")
            for name in set(names):
                GateFactory.registry[name] = subclass
            return subclass

        return wrapper

    @staticmethod
    def unregister(cls: type) -> None:
        for key in [k for k, v in GateFactory.registry.items() if v == cls]:
            del GateFactory.registry[key]


def _names(*names: str) -> List[str]:
    """A registry entry under each name and its lower-case form (the reference registers both spellings)."""
    out = []
    for n in names:
        out += [n, n.lower()]
    return list(dict.fromkeys(out))


# ---- decorators ------------------------------------------------------------------------------------------------------
class _Wrapped(Gate):
    def __init__(self, gate: Gate):
        self._gate = gate
        super().__init__(name=gate.name, num_wires=gate.num_wires, dim=gate.dimension, params=gate.params)

    def _validate_dimension(self, dim):
        self._gate._validate_dimension(dim)  # pylint: disable=protected-access


class Adjoint(_Wrapped):
    """Conjugate transpose of a gate (reference gate.py:278-298)."""

    def _data(self):
        return np.asarray(self._gate._data()).conj().T  # pylint: disable=protected-access


class Scale(_Wrapped):
    """A gate times a scalar (reference gate.py:301-323)."""

    def __init__(self, gate: Gate, scalar: float):
        self._scalar = scalar
        super().__init__(gate)

    def _data(self):
        return self._scalar * np.asarray(self._gate._data())  # pylint: disable=protected-access


# ---- continuous-variable gates in the Fock basis ---------------------------------------------------------------------
class FockGate(Gate):
    """A gate on ``num_wires`` bosonic modes truncated to ``cutoff`` Fock states (reference gate.py:331-348)."""

    def __init__(self, name: str, num_wires: int, cutoff: int, params: Optional[List[float]] = None):
        super().__init__(name=name, num_wires=num_wires, dim=cutoff, params=params)

    def _validate_dimension(self, dim):
        if dim < 2:
            raise ValueError("The dimension of a Fock gate must be greater than one.")


def _sqrt_factorials(n: int) -> np.ndarray:
    return np.sqrt(np.cumprod(np.concatenate(([1.0], np.arange(1, n, dtype=np.float64)))))


def _ladder_power_series(cutoff: int, modes: int, coeff: complex, raising: bool) -> np.ndarray:
    """exp(coeff * A) on `modes` modes truncated at `cutoff`, where A lowers (or raises) EVERY mode by one photon:
    A = a (one mode) or a b (two modes).  A is nilpotent on the truncated space, so the series is finite and the result
    is exact for every matrix element inside the cutoff.  Returns a (cutoff^modes x cutoff^modes) matrix."""
    a = np.diag(np.sqrt(np.arange(1, cutoff, dtype=np.float64)), k=1).astype(np.complex128)  # lowering operator
    step = a if modes == 1 else np.kron(a, a)
    if raising:
        step = step.conj().T
    out = np.eye(cutoff**modes, dtype=np.complex128)
    term = out.copy()
    for j in range(1, cutoff):
        term = term @ step * (coeff / j)
        out = out + term
    return out


@GateFactory.register(names=_names("Displacement", "D"))
class Displacement(FockGate):
    """D(alpha) = exp(alpha a^dag - conj(alpha) a), alpha = r e^{i phi} (reference gate.py:352-368; thewalrus
    ``displacement``).  Disentangled: D = e^{-|alpha|^2 / 2} exp(alpha a^dag) exp(-conj(alpha) a)."""

    def __init__(self, r: float, phi: float, cutoff: int = 2):
        super().__init__(name="Displacement", num_wires=1, cutoff=cutoff, params=[r, phi])

    def _data(self):
        r, phi = self.params
        alpha = r * cmath.exp(1j * phi)
        n = self.dimension
        return math.exp(-0.5 * r * r) * (_ladder_power_series(n, 1, alpha, True) @
                                          _ladder_power_series(n, 1, -alpha.conjugate(), False))


@GateFactory.register(names=_names("Squeezing"))
class Squeezing(FockGate):
    """S(z) = exp((conj(z) a^2 - z a^dag^2) / 2), z = r e^{i theta} (reference gate.py:372-388; thewalrus
    ``squeezing``).  Disentangled with tau = e^{i theta} tanh r:
    S = exp(-tau a^dag^2 / 2) (cosh r)^{-(a^dag a + 1/2)} exp(conj(tau) a^2 / 2)."""

    def __init__(self, r: float, theta: float, cutoff: int = 2):
        super().__init__(name="Squeezing", num_wires=1, cutoff=cutoff, params=[r, theta])

    def _data(self):
        r, theta = self.params
        n = self.dimension
        tau = cmath.exp(1j * theta) * math.tanh(r)
        a = np.diag(np.sqrt(np.arange(1, n, dtype=np.float64)), k=1).astype(np.complex128)
        a2 = a @ a

        def series(step, coeff):
            out = np.eye(n, dtype=np.complex128)
            term = out.copy()
            for j in range(1, n):
                term = term @ step * (coeff / j)
                out = out + term
            return out

        middle = np.diag(np.cosh(r) ** (-(np.arange(n) + 0.5))).astype(np.complex128)
        return series(a2.conj().T, -0.5 * tau) @ middle @ series(a2, 0.5 * tau.conjugate())


@GateFactory.register(names=_names("TwoModeSqueezing"))
class TwoModeSqueezing(FockGate):
    """S2(z) = exp(z a^dag b^dag - conj(z) a b), z = r e^{i theta} (reference gate.py:392-416; thewalrus
    ``two_mode_squeezing``).  Disentangled with tau = e^{i theta} tanh r:
    S2 = exp(tau a^dag b^dag) (cosh r)^{-(a^dag a + b^dag b + 1)} exp(-conj(tau) a b)."""

    def __init__(self, r: float, theta: float, cutoff: int = 2):
        super().__init__(name="TwoModeSqueezing", num_wires=2, cutoff=cutoff, params=[r, theta])

    def _data(self):
        r, theta = self.params
        n = self.dimension
        tau = cmath.exp(1j * theta) * math.tanh(r)
        photons = np.add.outer(np.arange(n), np.arange(n)).reshape(-1)
        middle = np.diag(np.cosh(r) ** (-(photons + 1.0))).astype(np.complex128)
        return _ladder_power_series(n, 2, tau, True) @ middle @ _ladder_power_series(n, 2, -tau.conjugate(), False)


@GateFactory.register(names=_names("Beamsplitter", "BS"))
class Beamsplitter(FockGate):
    """B(theta, phi) = exp(theta (e^{i phi} a b^dag - e^{-i phi} a^dag b)) (reference gate.py:419-441; thewalrus
    ``beamsplitter``).  Photon-number conserving: the creation operators transform linearly,
    a^dag -> cos(theta) a^dag - e^{-i phi} sin(theta) b^dag,  b^dag -> e^{i phi} sin(theta) a^dag + cos(theta) b^dag,
    and <m n| B |k l> follows from expanding (a^dag')^k (b^dag')^l |00> / sqrt(k! l!)."""

    def __init__(self, theta: float, phi: float, cutoff: int = 2):
        super().__init__(name="Beamsplitter", num_wires=2, cutoff=cutoff, params=[theta, phi])

    def _data(self):
        theta, phi = self.params
        n = self.dimension
        c, s = math.cos(theta), math.sin(theta)
        aa, ab = c, -cmath.exp(-1j * phi) * s  # a^dag' = aa a^dag + ab b^dag
        ba, bb = cmath.exp(1j * phi) * s, c    # b^dag' = ba a^dag + bb b^dag
        sf = _sqrt_factorials(2 * n)
        out = np.zeros((n, n, n, n), dtype=np.complex128)
        for k in range(n):
            for l in range(n):
                # (aa x + ab y)^k (ba x + bb y)^l: coefficient of x^m y^(k+l-m)
                for i in range(k + 1):          # i photons of the first factor go to mode a
                    for j in range(l + 1):      # j photons of the second factor go to mode a
                        m, q = i + j, k + l - i - j
                        if m >= n or q >= n:
                            continue
                        coeff = (math.comb(k, i) * aa**i * ab ** (k - i)) * (math.comb(l, j) * ba**j * bb ** (l - j))
                        out[m, q, k, l] += coeff * sf[m] * sf[q] / (sf[k] * sf[l])
        return out.reshape(n * n, n * n)


# ---- qubit gates -----------------------------------------------------------------------------------------------------
class QubitGate(Gate):
    """A gate on qubits (reference gate.py:444-458)."""

    def __init__(self, name: str, num_wires: int, params: Optional[List[float]] = None):
        super().__init__(name=name, num_wires=num_wires, dim=2, params=params)

    def _validate_dimension(self, dim):
        if dim != 2:
            raise ValueError("The dimension of a qubit gate must be exactly two.")


_R2 = 1 / math.sqrt(2)
_I2 = np.eye(2, dtype=np.complex128)


def _controlled(u: np.ndarray, controls: int = 1) -> np.ndarray:
    """|1..1><1..1| (x) u + (1 - |1..1><1..1|) (x) 1: the block of the last control pattern is u."""
    dim = u.shape[0] * 2**controls
    out = np.eye(dim, dtype=np.complex128)
    out[dim - u.shape[0]:, dim - u.shape[0]:] = u
    return out


def _rx(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)


def _ry(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def _rz(t):
    return np.diag([cmath.exp(-0.5j * t), cmath.exp(0.5j * t)]).astype(np.complex128)


def _phase(p):
    return np.diag([1, cmath.exp(1j * p)]).astype(np.complex128)


def _rot(phi, theta, omega):
    return _rz(omega) @ _ry(theta) @ _rz(phi)


_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.diag([1, -1]).astype(np.complex128)
_SWAP = np.eye(4, dtype=np.complex128)[[0, 2, 1, 3]]

_FIXED = {
    "Hadamard": (1, _R2 * np.array([[1, 1], [1, -1]], dtype=np.complex128)),
    "PauliX": (1, _X),
    "PauliY": (1, _Y),
    "PauliZ": (1, _Z),
    "S": (1, _phase(math.pi / 2)),
    "T": (1, _phase(math.pi / 4)),
    "SX": (1, 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=np.complex128)),
    "CX": (2, _controlled(_X)),
    "CY": (2, _controlled(_Y)),
    "CZ": (2, _controlled(_Z)),
    "SWAP": (2, _SWAP),
    "ISWAP": (2, np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)),
    "CSWAP": (3, _controlled(_SWAP)),
    "Toffoli": (3, _controlled(_X, controls=2)),
}


def _fixed_gate(cls_name: str, names: Sequence[str], doc: str) -> type:
    wires, matrix = _FIXED[cls_name]

    def __init__(self):
        QubitGate.__init__(self, name=cls_name, num_wires=wires)

    def _data(self):  # pylint: disable=unused-argument
        return matrix.copy()

    cls = type(cls_name, (QubitGate,), {"__init__": __init__, "_data": _data, "__doc__": doc, "__module__": __name__})
    return GateFactory.register(names=names)(cls)


Hadamard = _fixed_gate("Hadamard", _names("Hadamard", "H"), "Hadamard gate (reference gate.py:462-471).")
PauliX = _fixed_gate("PauliX", _names("PauliX", "X", "NOT"), "Pauli-X gate (reference gate.py:475-484).")
PauliY = _fixed_gate("PauliY", _names("PauliY", "Y"), "Pauli-Y gate (reference gate.py:488-497).")
PauliZ = _fixed_gate("PauliZ", _names("PauliZ", "Z"), "Pauli-Z gate (reference gate.py:501-510).")
S = _fixed_gate("S", _names("S"), "Phase gate diag(1, i) (reference gate.py:514-523).")
T = _fixed_gate("T", _names("T"), "pi/8 gate diag(1, e^{i pi/4}) (reference gate.py:527-536).")
SX = _fixed_gate("SX", _names("SX"), "Square root of X (reference gate.py:540-549).")
CX = _fixed_gate("CX", _names("CX", "CNOT"), "Controlled X, control on the first wire (reference gate.py:589-598).")
CY = _fixed_gate("CY", _names("CY"), "Controlled Y (reference gate.py:602-611).")
CZ = _fixed_gate("CZ", _names("CZ"), "Controlled Z (reference gate.py:615-624).")
SWAP = _fixed_gate("SWAP", _names("SWAP"), "Swap (reference gate.py:628-637).")
ISWAP = _fixed_gate("ISWAP", _names("ISWAP"), "iSWAP (reference gate.py:641-650).")
CSWAP = _fixed_gate("CSWAP", _names("CSWAP"), "Controlled swap (reference gate.py:654-672).")
Toffoli = _fixed_gate("Toffoli", _names("Toffoli"), "Doubly controlled X (reference gate.py:676-694).")


@GateFactory.register(names=_names("PhaseShift"))
class PhaseShift(QubitGate):
    """diag(1, e^{i phi}) (reference gate.py:553-567)."""

    def __init__(self, phi: float):
        super().__init__(name="PhaseShift", num_wires=1, params=[phi])

    def _data(self):
        return _phase(self.params[0])


@GateFactory.register(names=_names("CPhaseShift"))
class CPhaseShift(QubitGate):
    """Controlled phase shift (reference gate.py:571-585)."""

    def __init__(self, phi: float):
        super().__init__(name="CPhaseShift", num_wires=2, params=[phi])

    def _data(self):
        return _controlled(_phase(self.params[0]))


@GateFactory.register(names=_names("RX"))
class RX(QubitGate):
    """exp(-i theta X / 2) (reference gate.py:698-715)."""

    def __init__(self, theta: float):
        super().__init__(name="RX", num_wires=1, params=[theta])

    def _data(self):
        return _rx(self.params[0])


@GateFactory.register(names=_names("RY"))
class RY(QubitGate):
    """exp(-i theta Y / 2) (reference gate.py:719-737)."""

    def __init__(self, theta: float):
        super().__init__(name="RY", num_wires=1, params=[theta])

    def _data(self):
        return _ry(self.params[0])


@GateFactory.register(names=_names("RZ"))
class RZ(QubitGate):
    """exp(-i theta Z / 2) (reference gate.py:741-757)."""

    def __init__(self, theta: float):
        super().__init__(name="RZ", num_wires=1, params=[theta])

    def _data(self):
        return _rz(self.params[0])


@GateFactory.register(names=_names("Rot"))
class Rot(QubitGate):
    """RZ(omega) RY(theta) RZ(phi) (reference gate.py:761-793)."""

    def __init__(self, phi: float, theta: float, omega: float):
        super().__init__(name="Rot", num_wires=1, params=[phi, theta, omega])

    def _data(self):
        return _rot(*self.params)


@GateFactory.register(names=_names("CRX"))
class CRX(QubitGate):
    """Controlled RX (reference gate.py:797-814)."""

    def __init__(self, theta: float):
        super().__init__(name="CRX", num_wires=2, params=[theta])

    def _data(self):
        return _controlled(_rx(self.params[0]))


@GateFactory.register(names=_names("CRY"))
class CRY(QubitGate):
    """Controlled RY (reference gate.py:818-835)."""

    def __init__(self, theta: float):
        super().__init__(name="CRY", num_wires=2, params=[theta])

    def _data(self):
        return _controlled(_ry(self.params[0]))


@GateFactory.register(names=_names("CRZ"))
class CRZ(QubitGate):
    """Controlled RZ (reference gate.py:839-858)."""

    def __init__(self, theta: float):
        super().__init__(name="CRZ", num_wires=2, params=[theta])

    def _data(self):
        return _controlled(_rz(self.params[0]))


@GateFactory.register(names=_names("CRot"))
class CRot(QubitGate):
    """Controlled Rot (reference gate.py:862-886)."""

    def __init__(self, phi: float, theta: float, omega: float):
        super().__init__(name="CRot", num_wires=2, params=[phi, theta, omega])

    def _data(self):
        return _controlled(_rot(*self.params))


@GateFactory.register(names=_names("U1"))
class U1(QubitGate):
    """diag(1, e^{i phi}) (reference gate.py:890-904)."""

    def __init__(self, phi: float):
        super().__init__(name="U1", num_wires=1, params=[phi])

    def _data(self):
        return _phase(self.params[0])


@GateFactory.register(names=_names("U2"))
class U2(QubitGate):
    """(1 / sqrt 2) [[1, -e^{i lam}], [e^{i phi}, e^{i (phi + lam)}]] (reference gate.py:908-926)."""

    def __init__(self, phi: float, lam: float):
        super().__init__(name="U2", num_wires=1, params=[phi, lam])

    def _data(self):
        phi, lam = self.params
        return _R2 * np.array([[1, -cmath.exp(1j * lam)], [cmath.exp(1j * phi), cmath.exp(1j * (phi + lam))]],
                              dtype=np.complex128)


@GateFactory.register(names=_names("U3"))
class U3(QubitGate):
    """[[c, -e^{i lam} s], [e^{i phi} s, e^{i (phi + lam)} c]], c = cos(theta / 2), s = sin(theta / 2)
    (reference gate.py:930-952)."""

    def __init__(self, theta: float, phi: float, lam: float):
        super().__init__(name="U3", num_wires=1, params=[theta, phi, lam])

    def _data(self):
        theta, phi, lam = self.params
        c, s = math.cos(theta / 2), math.sin(theta / 2)
        return np.array([[c, -cmath.exp(1j * lam) * s], [cmath.exp(1j * phi) * s, cmath.exp(1j * (phi + lam)) * c]],
                        dtype=np.complex128)
