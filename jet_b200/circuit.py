"""Circuits as tensor networks (reference python/jet/circuit.py:18-215: ``Operation``, ``Wire``, ``Circuit``), restated
over the B200 engine.

The bookkeeping is the reference's: wire ``i`` carries the index label ``"i-depth"``; a gate consumes the current
labels of its wires as inputs and advances their depth for its outputs (gate tensor indices = outputs ++ inputs); a
state closes its wires; ``take_expected_value`` appends the observable, the adjoints of every gate in reverse order and
closing vacuum states.  ``tensor_network`` returns the engine's ``TensorNetwork`` (same tensors, same order, as the
reference's).  Two additions hand a circuit to the plan engine without going through node-by-node host calls:
``network_file`` (leaves + a contraction path as a ``NetworkFile``) and ``amplitude`` (one fused, graph-captured
contraction on the GPU: ``ContractionPlan``)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List, Optional, Sequence, Union

import numpy as np

from .gate import Adjoint, Gate
from .state import Qudit, State

__all__ = ["Circuit", "Operation", "Wire"]


@dataclass(frozen=True)
class Operation:
    """A gate or state together with the wires it is attached to (reference circuit.py:18-29)."""

    part: Union[Gate, State]
    wire_ids: Sequence[int]


@dataclass
class Wire:
    """The chain of tensor indices that follows one qudit through the circuit (reference circuit.py:32-49)."""

    id_: int
    depth: int = 0
    closed: bool = False

    @property
    def index(self) -> str:
        return f"{self.id_}-{self.depth}"


class Circuit:
    """``num_wires`` qudits of dimension ``dim``, each starting in the vacuum state (reference circuit.py:52-215)."""

    def __init__(self, num_wires: int, dim: int = 2):
        self._dim = dim
        self._wires: List[Wire] = [Wire(i) for i in range(num_wires)]
        self._ops: List[Operation] = []
        for wire in self._wires:
            state = Qudit(dim=dim)
            state.indices = [wire.index]
            self._ops.append(Operation(part=state, wire_ids=[wire.id_]))

    @property
    def dimension(self) -> int:
        return self._dim

    @property
    def operations(self) -> Iterator[Operation]:
        """The initial qudits (one per wire, in wire order), then everything appended, in order."""
        return iter(self._ops)

    @property
    def wires(self) -> Iterator[Wire]:
        return iter(self._wires)

    def indices(self, wire_ids: Iterator[int]) -> Iterator[str]:
        """Current index label of each listed wire."""
        return (self._wires[i].index for i in wire_ids)

    def _validate_wire_ids(self, wire_ids: Sequence[int]) -> None:
        n = len(self._wires)
        for wire_id in wire_ids:
            if not 0 <= wire_id < n:
                raise ValueError(f"Wire ID {wire_id} falls outside the range [0, {n}).")
            if list(wire_ids).count(wire_id) > 1:
                raise ValueError(f"Wire ID {wire_id} is specified more than once.")
            if self._wires[wire_id].closed:
                raise ValueError(f"Wire {wire_id} is closed.")

    def append_gate(self, gate: Gate, wire_ids: Sequence[int]) -> None:
        """Applies ``gate`` to the listed wires."""
        self._validate_wire_ids(wire_ids)
        if len(wire_ids) != gate.num_wires:
            raise ValueError(f"Number of wire IDs ({len(wire_ids)}) must match the number of "
                             f"wires connected to the gate ({gate.num_wires}).")
        inputs = list(self.indices(wire_ids))
        for i in wire_ids:
            self._wires[i].depth += 1
        gate.indices = list(self.indices(wire_ids)) + inputs
        self._ops.append(Operation(part=gate, wire_ids=wire_ids))

    def append_state(self, state: State, wire_ids: Sequence[int]) -> None:
        """Terminates the listed wires with ``state``."""
        self._validate_wire_ids(wire_ids)
        if len(wire_ids) != state.num_wires:
            raise ValueError(f"Number of wire IDs ({len(wire_ids)}) must match the number of "
                             f"wires connected to the state ({state.num_wires}).")
        for i in wire_ids:
            self._wires[i].closed = True
        state.indices = list(self.indices(wire_ids))
        self._ops.append(Operation(part=state, wire_ids=wire_ids))

    def take_expected_value(self, observable: Iterator[Operation]) -> None:
        """Closes the circuit into <0| U^dag O U |0>: the observable's gates, then the adjoint of every gate applied so
        far in reverse order, then a vacuum state on every wire.  Nothing can be appended afterwards."""
        first_gate, end = len(self._wires), len(self._ops)
        for op in observable:
            self.append_gate(gate=op.part, wire_ids=op.wire_ids)
        for op in reversed(self._ops[first_gate:end]):
            self.append_gate(gate=Adjoint(gate=op.part), wire_ids=op.wire_ids)
        for op in reversed(self._ops[:first_gate]):  # the initial qudits are real: no adjoint needed
            self.append_state(state=Qudit(dim=self.dimension), wire_ids=op.wire_ids)

    def tensor_network(self, dtype: np.dtype = np.complex128):
        """The circuit as the engine's ``TensorNetwork`` (one tensor per operation, in order)."""
        from .jet import TensorNetwork  # the compiled bindings are only needed here

        tn = TensorNetwork(dtype=dtype)
        for op in self._ops:
            tn.add_tensor(op.part.tensor(dtype=dtype))
        return tn

    # ---- additions: straight to the plan engine ----------------------------------------------------------------------
    def leaves(self, dtype: np.dtype = np.complex128):
        """[(index labels, ndarray)] of every operation, in order — the leaves of the tensor network."""
        out = []
        for op in self._ops:
            part = op.part
            data = np.asarray(part._data(), dtype=dtype).reshape(-1)  # pylint: disable=protected-access
            default = 2 * part.num_wires if isinstance(part, Gate) else part.num_wires
            indices = list(part.indices) if part.indices is not None else [str(i) for i in range(default)]
            dim = int(round(len(data) ** (1.0 / len(indices))))
            out.append((indices, data.reshape([dim] * len(indices))))
        return out

    def network_file(self, dtype: np.dtype = np.complex128, path: Optional[Sequence[Sequence[int]]] = None, trials: int = 8):
        """The circuit as a ``NetworkFile`` (the reference's JSON model: leaves + path).  Without ``path`` a pairwise
        contraction path is searched (``jet_b200.pathfinder.search``: seeded greedy, best of ``trials``)."""
        from .pathfinder import search
        from .plan import NetworkFile

        leaves = self.leaves(dtype)
        if path is None:
            dims = {}
            for idx, arr in leaves:
                for i, d in zip(idx, arr.shape):
                    dims[i] = int(d)
            path, _ = search([idx for idx, _ in leaves], dims, trials=trials)
        return NetworkFile(leaves, path)

    def amplitude(self, dtype: np.dtype = np.complex128, sliced: Sequence[str] = (), path=None, device: int = 0):
        """Contracts the whole circuit on the GPU as one plan (fused chains, CUDA graph, FP64 sum over the slices of
        ``sliced``).  A closed circuit gives a complex scalar; a circuit with open wires gives ``(index labels, ndarray)``
        in the engine's result order (``left ++ right`` of the last contraction, as ``ContractTensors`` would)."""
        from .plan import ContractionPlan

        net = self.network_file(dtype, path)
        with ContractionPlan(net, list(sliced), device=device) as plan:
            out = np.asarray(plan.amplitude())
            labels = list(plan.result_indices)
        return out.reshape(-1)[0] if not labels else (labels, out)
