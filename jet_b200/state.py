"""Quantum states as tensors (reference python/jet/state.py:17-153: ``State``, ``Qudit``, ``QuditRegister``, ``Qubit``,
``QubitRegister``), restated over the B200 engine's ``Tensor`` factory.  A state on n wires of dimension d is a rank-n
tensor of shape ``[d] * n``; the default is the vacuum |0...0>."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Optional, Sequence

import numpy as np

from .gate import _check_indices

__all__ = ["State", "Qudit", "QuditRegister", "Qubit", "QubitRegister"]


class State(ABC):
    """Base of the state classes (reference state.py:17-93)."""

    def __init__(self, name: str, num_wires: int):
        self.name = name
        self._indices = None
        self._num_wires = num_wires

    @property
    def indices(self) -> Optional[List[str]]:
        return self._indices

    @indices.setter
    def indices(self, indices: Optional[Sequence[str]]) -> None:
        _check_indices(indices, self._num_wires,
                       "States must have one index per wire. Received {} indices for " + f"{self._num_wires} wires.")
        self._indices = indices

    @property
    def num_wires(self) -> int:
        return self._num_wires

    def __eq__(self, other) -> bool:
        return bool(np.all(self._data() == other._data()))  # pylint: disable=protected-access

    def __ne__(self, other) -> bool:
        return not self == other

    __hash__ = None

    @abstractmethod
    def _data(self) -> np.ndarray:
        """The state vector (row-major over the wires)."""

    def tensor(self, dtype: np.dtype = np.complex128):
        """The state as an engine tensor with labels ``indices`` (default "0", "1", ...)."""
        from .jet import Tensor  # the compiled bindings are only needed here

        data = np.asarray(self._data()).reshape(-1)
        indices = list(self.indices) if self.indices is not None else [str(i) for i in range(self._num_wires)]
        dim = int(round(len(data) ** (1.0 / len(indices))))
        return Tensor(indices=indices, shape=[dim] * len(indices), data=data, dtype=dtype)


class Qudit(State):
    """One qudit of dimension ``dim``; ``data`` = its state vector, default |0> (reference state.py:96-112)."""

    def __init__(self, dim: int, data: Optional[np.ndarray] = None):
        name = "Qubit" if dim == 2 else f"Qudit(d={dim})"
        super().__init__(name=name, num_wires=1)
        self._vector = _vacuum(dim) if data is None else np.asarray(data).reshape(-1)

    def _data(self) -> np.ndarray:
        return self._vector


class QuditRegister(State):
    """``size`` qudits of dimension ``dim``; ``data`` = the joint state vector, default |0...0>
    (reference state.py:115-132)."""

    def __init__(self, dim: int, size: int, data: Optional[np.ndarray] = None):
        name = f"Qubit[{size}]" if dim == 2 else f"Qudit(d={dim})[{size}]"
        super().__init__(name=name, num_wires=size)
        self._vector = _vacuum(dim**size) if data is None else np.asarray(data).reshape(-1)

    def _data(self) -> np.ndarray:
        return self._vector


def _vacuum(n: int) -> np.ndarray:
    v = np.zeros(n, dtype=np.complex128)
    v[0] = 1
    return v


def Qubit(data: Optional[np.ndarray] = None) -> Qudit:
    """A qudit of dimension two (reference state.py:135-142)."""
    return Qudit(dim=2, data=data)


def QubitRegister(size: int, data: Optional[np.ndarray] = None) -> QuditRegister:
    """A register of qubits (reference state.py:145-153)."""
    return QuditRegister(dim=2, size=size, data=data)
