"""What the reference's XIR interpreter computes once a program has been turned into a ``Circuit``
(reference python/jet/interpreter.py:437-530: ``_compute_amplitude``, ``_compute_probabilities``,
``_compute_expected_value``, ``_simulate``), on the plan engine: the circuit's tensor network goes to the GPU as ONE
``ContractionPlan`` (searched path, fused chains, CUDA graph) instead of the reference's
``TaskBasedContractor`` over a randomly sampled path.  The XIR parsing layer itself (``run_xir_program``) needs the
``xir`` package, which this image does not have; a caller with ``xir`` maps statements to these functions exactly as
the reference does (interpreter.py:120-226).
"""
from __future__ import annotations

from copy import deepcopy
from typing import Iterator, List, Sequence, Tuple

import numpy as np

from .circuit import Circuit, Operation
from .state import Qudit

__all__ = ["simulate", "compute_amplitude", "compute_probabilities", "compute_expected_value"]


def simulate(circuit: Circuit, dtype: np.dtype = np.complex128, device: int = 0) -> Tuple[List[str], np.ndarray]:
    """Contracts the circuit's tensor network.  Returns (index labels, ndarray): no labels and a 0-d array for a closed
    circuit, the labels of the open wires (engine order) and the output tensor otherwise.  A network of one tensor
    needs no contraction (interpreter.py:513-516)."""
    leaves = circuit.leaves(dtype)
    if len(leaves) == 1:
        return list(leaves[0][0]), np.asarray(leaves[0][1], dtype=np.complex128)
    out = circuit.amplitude(dtype=dtype, device=device)
    if isinstance(out, tuple):
        return list(out[0]), np.asarray(out[1])
    return [], np.asarray(out)


def compute_amplitude(circuit: Circuit, state: Sequence[int], dtype: np.dtype = np.complex128, device: int = 0):
    """<state| U |0...0>: every wire i is closed with the basis state ``state[i]`` (interpreter.py:437-460).  The
    circuit passed in is not modified."""
    wires = sum(1 for _ in circuit.wires)
    if len(state) != wires:
        raise ValueError(f"The state has {len(state)} (!= {wires}) entries.")
    closed = deepcopy(circuit)
    for i, value in enumerate(state):
        if not 0 <= value < circuit.dimension:
            raise ValueError(f"State entry {value} falls outside the range [0, {circuit.dimension}).")
        data = np.zeros(circuit.dimension, dtype=np.complex128)
        data[value] = 1
        closed.append_state(Qudit(dim=circuit.dimension, data=data), wire_ids=[i])
    _, value = simulate(closed, dtype, device)
    return np.dtype(dtype).type(value.reshape(-1)[0])


def compute_probabilities(circuit: Circuit, dtype: np.dtype = np.complex128, device: int = 0) -> np.ndarray:
    """|amplitude|^2 of every basis state, wire 0 slowest (interpreter.py:463-479)."""
    labels, tensor = simulate(circuit, dtype, device)
    want = [wire.index for wire in circuit.wires]
    if sorted(labels) != sorted(want):
        raise ValueError("Probabilities need a circuit whose wires are all open.")
    amplitudes = np.transpose(tensor, [labels.index(i) for i in want]).reshape(-1)
    return (amplitudes.conj() * amplitudes).astype(dtype)


def compute_expected_value(circuit: Circuit, observable: Iterator[Operation], dtype: np.dtype = np.complex128,
                           device: int = 0):
    """<0| U^dag O U |0> through ``Circuit.take_expected_value`` on a copy (interpreter.py:482-501)."""
    closed = deepcopy(circuit)
    closed.take_expected_value(observable)
    _, value = simulate(closed, dtype, device)
    return np.dtype(dtype).type(value.reshape(-1)[0])
