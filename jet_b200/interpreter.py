"""Runs XIR programs on the engine (reference python/jet/interpreter.py:29-435: ``get_xir_manifest``,
``run_xir_program``).  Same statement semantics, validation order and error messages as the reference; the differences:

* programs come from ``jet_b200.xir_lite.parse_script`` (the ``xir`` package is absent here; ``run_xir_script(text)``
  parses and runs in one call);
* every output statement (``amplitude``, ``probabilities``, ``expval``) contracts the circuit built so far as ONE GPU plan
  (``jet_b200.simulate``) instead of a ``TaskBasedContractor`` over a randomly sampled path.
"""
from __future__ import annotations

import warnings
from inspect import signature
from typing import Any, Callable, Dict, Iterator, List, Set, Union

import numpy as np

from .circuit import Circuit, Operation
from .gate import FockGate, GateFactory
from .simulate import compute_amplitude, compute_expected_value, compute_probabilities
from .xir_lite import Declaration, Program, Statement, parse_script

__all__ = ["get_xir_manifest", "run_xir_program", "run_xir_script"]

Params = Dict[str, Any]
Wires = Dict[Any, Any]
Stack = Set[str]
StatementGenerator = Callable[[Params, Wires, Stack], Iterator[Statement]]


def _get_xir_outputs() -> Iterator[str]:
    yield from sorted(("Amplitude", "amplitude", "Expval", "expval", "Probabilities", "probabilities"))


def get_xir_manifest() -> Program:
    """Declarations of every registered gate (parameters = the constructor's required arguments, wires 0 .. n-1) and of
    the supported outputs (reference interpreter.py:29-68)."""
    program = Program()
    for name, cls in sorted(GateFactory.registry.items()):
        keys = [p.name for p in signature(cls.__init__).parameters.values() if p.default is p.empty][1:]
        gate = cls(*[None for _ in keys])
        program.add_declaration(Declaration("gate", name, keys, tuple(range(gate.num_wires))))
    for name in _get_xir_outputs():
        program.add_declaration(Declaration("out", name))
    return program


def run_xir_script(text: str, eval_pi: bool = True) -> List[Union[np.number, np.ndarray]]:
    """``run_xir_program(parse_script(text))``."""
    return run_xir_program(parse_script(text, eval_pi=eval_pi))


def run_xir_program(program: Program) -> List[Union[np.number, np.ndarray]]:
    """Executes the program: gate statements build a circuit over the wires the program uses; each output statement
    appends one value (reference interpreter.py:77-226)."""
    result: List[Union[np.number, np.ndarray]] = []
    program = Program.merge(get_xir_manifest(), program)
    _validate_options(program)
    num_wires = len(program.wires)
    dimension = program.options.get("dimension", 2)
    circuit = Circuit(num_wires=num_wires, dim=dimension)
    everything = tuple(range(num_wires))

    for stmt in _resolve_statements(program):
        if stmt.name in GateFactory.registry:
            gate = GateFactory.create(stmt.name, **stmt.params)
            if isinstance(gate, FockGate):
                gate.dimension = circuit.dimension
            if gate.dimension != circuit.dimension:
                raise ValueError(f"Statement '{stmt}' applies a gate with a dimension ({gate.dimension}) "
                                 f"that differs from the dimension of the circuit ({circuit.dimension}).")
            circuit.append_gate(gate, wire_ids=list(stmt.wires))

        elif stmt.name in ("Amplitude", "amplitude"):
            if not isinstance(stmt.params, dict) or "state" not in stmt.params:
                raise ValueError(f"Statement '{stmt}' is missing a 'state' parameter.")
            state = stmt.params["state"]
            if not isinstance(state, list):
                raise ValueError(f"Statement '{stmt}' has a 'state' parameter which is not an array.")
            if not all(0 <= entry < dimension for entry in state):
                raise ValueError(f"Statement '{stmt}' has a 'state' parameter with at least "
                                 f"one entry that falls outside the range [0, {dimension}).")
            if len(state) != num_wires:
                raise ValueError(f"Statement '{stmt}' has a 'state' parameter with "
                                 f"{len(state)} (!= {num_wires}) entries.")
            if tuple(stmt.wires) != everything:
                raise ValueError(f"Statement '{stmt}' must be applied to [0 .. {num_wires - 1}].")
            result.append(compute_amplitude(circuit, state))

        elif stmt.name in ("Probabilities", "probabilities"):
            if tuple(stmt.wires) != everything:
                raise ValueError(f"Statement '{stmt}' must be applied to [0 .. {num_wires - 1}].")
            result.append(compute_probabilities(circuit))

        elif stmt.name in ("Expval", "expval"):
            if not isinstance(stmt.params, dict) or "observable" not in stmt.params:
                raise ValueError(f"Statement '{stmt}' is missing an 'observable' parameter.")
            observable = stmt.params["observable"]
            if observable not in program.observables:
                raise ValueError(f"Statement '{stmt}' has an 'observable' parameter which "
                                 f"references an undefined observable.")
            if program.search("obs", "params", observable):
                raise ValueError(f"Statement '{stmt}' has an 'observable' parameter which "
                                 f"references a parameterized observable.")
            have, want = tuple(stmt.wires), program.search("obs", "wires", observable)
            if len(have) != len(want):
                raise ValueError(f"Statement '{stmt}' has an 'observable' parameter which "
                                 f"applies the wrong number of wires.")
            wires = {label: have[i] for i, label in enumerate(want)}
            operations = list(_operations_of_observable(program.observables[observable], wires))
            result.append(compute_expected_value(circuit, operations))

        else:
            raise ValueError(f"Statement '{stmt}' is not supported.")
    return result


def _validate_options(program: Program) -> None:
    """reference interpreter.py:230-252"""
    for option, value in program.options.items():
        if option == "dimension":
            if not isinstance(value, int) or isinstance(value, bool):
                raise ValueError("Option 'dimension' must be an integer.")
            if value < 2:
                raise ValueError("Option 'dimension' must be greater than one.")
        else:
            warnings.warn(f"Option '{option}' is not supported and will be ignored.")


def _resolve_statements(program: Program) -> Iterator[Statement]:
    """Top-level statements with every user-defined (composite) gate expanded into registered gates, parameters and
    wires substituted (reference interpreter.py:255-350).  Statements that apply no known gate pass through."""
    signatures = {decl.name: {"params": decl.params, "wires": decl.wires, "statements": program.gates.get(decl.name)}
                  for decl in program.declarations["gate"]}
    generators: Dict[str, StatementGenerator] = {}

    def terminal(name: str) -> StatementGenerator:
        def generator(params: Params, wires: Wires, _: Stack) -> Iterator[Statement]:
            yield Statement(name, params, tuple(wires[i] for i in sorted(wires, key=int)))

        return generator

    def composite(name: str) -> StatementGenerator:
        def generator(params: Params, wires: Wires, stack: Stack) -> Iterator[Statement]:
            if name in stack:
                raise ValueError(f"Gate '{name}' has a circular dependency.")
            for stmt in signatures[name]["statements"]:
                stmt_params = _bind_params(signatures, stmt)
                stmt_wires = _bind_wires(signatures, stmt)
                eval_params = {key: params.get(val, val) if isinstance(val, str) else val for key, val in stmt_params.items()}
                eval_wires = {key: wires.get(val, val) for key, val in stmt_wires.items()}
                yield from generators[stmt.name](eval_params, eval_wires, stack | {name})

        return generator

    for name in GateFactory.registry:
        generators[name] = terminal(name)
    for name in program.gates:
        if signatures[name]["statements"] is not None:
            generators[name] = composite(name)
    for stmt in program.statements:
        if stmt.name in generators:
            yield from generators[stmt.name](_bind_params(signatures, stmt), _bind_wires(signatures, stmt), set())
        else:
            yield stmt


def _bind_params(signatures, stmt: Statement) -> Params:
    if stmt.name not in signatures:
        raise ValueError(f"Statement '{stmt}' applies a gate which has not been defined.")
    have, want = stmt.params, signatures[stmt.name]["params"]
    if isinstance(have, list):
        if len(have) != len(want):
            raise ValueError(f"Statement '{stmt}' has the wrong number of parameters.")
        return {name: have[i] for i, name in enumerate(want)}
    if set(have) != set(want):
        raise ValueError(f"Statement '{stmt}' has an invalid set of parameters.")
    return {name: have[name] for name in want}


def _bind_wires(signatures, stmt: Statement) -> Wires:
    if stmt.name not in signatures:
        raise ValueError(f"Statement '{stmt}' applies a gate which has not been defined.")
    have, want = stmt.wires, signatures[stmt.name]["wires"]
    if len(have) != len(want):
        raise ValueError(f"Statement '{stmt}' has the wrong number of wires.")
    return {name: have[i] for i, name in enumerate(want)}


def _operations_of_observable(stmts, wires: Wires) -> Iterator[Operation]:
    """One scaled gate per factor of every term (reference interpreter.py:407-434)."""
    for stmt in stmts:
        try:
            scalar = float(stmt.pref)
        except (TypeError, ValueError) as exc:
            raise ValueError(f"Observable statement '{stmt}' has a prefactor ({stmt.pref}) "
                             f"which cannot be converted to a floating-point number.") from exc
        for gate_name, wire_label in stmt.terms:
            yield Operation(part=GateFactory.create(gate_name, scalar=scalar), wire_ids=[wires[wire_label]])
