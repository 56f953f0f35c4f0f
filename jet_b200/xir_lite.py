"""A small reader for XIR scripts — the subset the reference's interpreter consumes (reference
python/jet/interpreter.py:77-435 works on ``xir.Program`` objects from the ``xir`` package, which this image does not
have).  ``parse_script(text)`` returns a ``Program`` with the attributes that interpreter reads: ``options``,
``statements``, ``wires``, ``declarations["gate" | "obs" | "out"]``, ``gates`` (definitions), ``observables``.

Grammar covered (everything the reference's interpreter tests use, python/tests/test_interpreter.py)::

    use <name>;                                   ignored
    options: key: value; ... end;
    gate NAME [(p, q)] [[w0, w1]] ;               declaration
    gate NAME [(p, q)] [[w0, w1]] : stmt; ... end;   definition (statements may use parameter / wire names)
    obs  NAME [(p)] [[w0]] ;                      declaration
    obs  NAME [(p)] [[w0]] : pref, OP[w] [@ OP[w]]; ... end;
    out  NAME;
    NAME [(v, ...) | (key: v, ...)] | [w, ...];   application (gates and outputs alike)

Values: integers, decimals, ``pi`` (kept as the symbol ``PI`` unless ``eval_pi``), arithmetic on those with
``+ - * /`` and parentheses, identifiers, ``true`` / ``false``, arrays ``[v, ...]``.  ``str(statement)`` reproduces
the script form (the reference's error messages quote it).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

__all__ = ["parse_script", "Program", "Statement", "Declaration", "ObservableStmt"]

_TOKEN = re.compile(r"\s*(?:(//[^\n]*)|(\d+\.\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|\d+(?:[eE][-+]?\d+)?)|"
                    r"([A-Za-z_][A-Za-z_0-9]*)|(<[^>]*>)|(.))")


class Symbol(str):
    """An identifier or unevaluated expression inside a parameter list (printed verbatim)."""


def _show(value: Any) -> str:
    if isinstance(value, list):
        return "[" + ", ".join(_show(v) for v in value) + "]"
    if isinstance(value, bool):
        return "true" if value else "false"
    return str(value)


@dataclass
class Statement:
    """``name(params) | [wires]``: ``params`` is a list (positional) or a dict (``key: value``)."""

    name: str
    params: Union[List[Any], Dict[str, Any]]
    wires: Tuple[Any, ...]

    def __str__(self) -> str:
        if isinstance(self.params, dict):
            inner = ", ".join(f"{k}: {_show(v)}" for k, v in self.params.items())
        else:
            inner = ", ".join(_show(v) for v in self.params)
        head = f"{self.name}({inner})" if inner else self.name
        return f"{head} | [" + ", ".join(str(w) for w in self.wires) + "]"


@dataclass
class Declaration:
    type_: str
    name: str
    params: List[str] = field(default_factory=list)
    wires: Tuple[Any, ...] = ()

    def __str__(self) -> str:
        text = f"{self.type_} {self.name}"
        if self.params:
            text += "(" + ", ".join(self.params) + ")"
        if self.wires:
            text += "[" + ", ".join(str(w) for w in self.wires) + "]"
        return text


@dataclass
class ObservableStmt:
    """``pref, OP[w] @ OP[w]``."""

    pref: Any
    terms: List[Tuple[str, Any]]

    def __str__(self) -> str:
        return f"{_show(self.pref)}, " + " @ ".join(f"{name}[{wire}]" for name, wire in self.terms)


class Program:
    def __init__(self):
        self.options: Dict[str, Any] = {}
        self.statements: List[Statement] = []
        self.declarations: Dict[str, List[Declaration]] = {"gate": [], "obs": [], "out": [], "func": []}
        self.gates: Dict[str, List[Statement]] = {}
        self.observables: Dict[str, List[ObservableStmt]] = {}
        self.includes: List[str] = []

    @property
    def wires(self) -> List[Any]:
        """Wires the top-level statements touch, in increasing order."""
        seen = []
        for stmt in self.statements:
            for w in stmt.wires:
                if w not in seen:
                    seen.append(w)
        return sorted(seen, key=lambda w: (isinstance(w, str), w))

    def add_declaration(self, decl: Declaration) -> None:
        self.declarations[decl.type_] = [d for d in self.declarations[decl.type_] if d.name != decl.name] + [decl]

    def search(self, type_: str, attr: str, name: str):
        """``params`` or ``wires`` of the declaration ``name`` of kind ``type_`` (KeyError-free: () when absent)."""
        for decl in self.declarations.get(type_, []):
            if decl.name == name:
                return list(decl.params) if attr == "params" else tuple(decl.wires)
        return [] if attr == "params" else ()

    @staticmethod
    def merge(*programs: "Program") -> "Program":
        """Declarations, definitions and options of every program (later ones win), statements concatenated."""
        out = Program()
        for p in programs:
            out.options.update(p.options)
            out.statements += p.statements
            out.includes += p.includes
            for kind, decls in p.declarations.items():
                for d in decls:
                    out.add_declaration(d)
            out.gates.update(p.gates)
            out.observables.update(p.observables)
        return out

    def serialize(self, minimize: bool = False) -> str:
        """Declarations only (what a manifest holds), ``gate`` before ``out``, in insertion order."""
        lines = [str(d) + ";" for kind in ("gate", "obs", "out") for d in self.declarations[kind]]
        return (" " if minimize else "\n").join(lines)


class _Parser:
    def __init__(self, text: str, eval_pi: bool):
        self.tokens: List[Tuple[str, str]] = []
        pos = 0
        while pos < len(text):
            m = _TOKEN.match(text, pos)
            if m is None:
                break
            pos = m.end()
            if m.group(1) is not None:
                continue
            if m.group(2) is not None:
                self.tokens.append(("num", m.group(2)))
            elif m.group(3) is not None:
                self.tokens.append(("id", m.group(3)))
            elif m.group(4) is not None:
                self.tokens.append(("inc", m.group(4)))
            elif m.group(5) is not None and not m.group(5).isspace():
                self.tokens.append(("op", m.group(5)))
        self.i = 0
        self.eval_pi = eval_pi

    # -- token helpers
    def peek(self, k: int = 0) -> Tuple[str, str]:
        return self.tokens[self.i + k] if self.i + k < len(self.tokens) else ("eof", "")

    def take(self) -> Tuple[str, str]:
        tok = self.peek()
        self.i += 1
        return tok

    def accept(self, value: str) -> bool:
        if self.peek()[1] == value and self.peek()[0] in ("op", "id"):
            self.i += 1
            return True
        return False

    def expect(self, value: str) -> None:
        if not self.accept(value):
            raise ValueError(f"XIR syntax error: expected '{value}' but found '{self.peek()[1]}'.")

    def name(self) -> str:
        kind, value = self.take()
        if kind != "id":
            raise ValueError(f"XIR syntax error: expected a name but found '{value}'.")
        return value

    # -- values
    def value(self) -> Any:
        if self.peek() == ("op", "["):
            return self.array()
        return self.expression()

    def array(self) -> List[Any]:
        self.expect("[")
        items = []
        if not self.accept("]"):
            while True:
                items.append(self.value())
                if self.accept("]"):
                    break
                self.expect(",")
        return items

    def expression(self) -> Any:
        left = self.term()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.take()[1]
            left = self.combine(left, op, self.term())
        return left

    def term(self) -> Any:
        left = self.factor()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.take()[1]
            left = self.combine(left, op, self.factor())
        return left

    def factor(self) -> Any:
        kind, value = self.take()
        if (kind, value) == ("op", "-"):
            inner = self.factor()
            return -inner if isinstance(inner, (int, float)) else Symbol(f"-{inner}")
        if (kind, value) == ("op", "("):
            inner = self.expression()
            self.expect(")")
            return inner
        if kind == "num":
            return int(value) if re.fullmatch(r"\d+", value) else float(value)
        if kind == "id":
            if value in ("pi", "PI"):
                return math.pi if self.eval_pi else Symbol("PI")
            if value in ("true", "false"):
                return value == "true"
            return Symbol(value)
        raise ValueError(f"XIR syntax error: unexpected '{value}' in an expression.")

    @staticmethod
    def combine(a: Any, op: str, b: Any) -> Any:
        if isinstance(a, (int, float)) and isinstance(b, (int, float)) and not isinstance(a, bool) and not isinstance(b, bool):
            if op == "+":
                return a + b
            if op == "-":
                return a - b
            if op == "*":
                return a * b
            return a / b
        return Symbol(f"{a}{op}{b}")

    # -- pieces of statements
    def params(self) -> Union[List[Any], Dict[str, Any]]:
        """``( ... )`` of an application: positional values or ``key: value`` pairs; [] when absent."""
        if not self.accept("("):
            return []
        if self.accept(")"):
            return []
        if self.peek()[0] == "id" and self.peek(1) == ("op", ":"):
            out: Dict[str, Any] = {}
            while True:
                key = self.name()
                self.expect(":")
                out[key] = self.value()
                if self.accept(")"):
                    return out
                self.expect(",")
        items = []
        while True:
            items.append(self.value())
            if self.accept(")"):
                return items
            self.expect(",")

    def wire_list(self) -> Tuple[Any, ...]:
        self.expect("[")
        wires: List[Any] = []
        if not self.accept("]"):
            while True:
                kind, value = self.take()
                if kind == "num" and re.fullmatch(r"\d+", value):
                    wires.append(int(value))
                elif kind == "id":
                    wires.append(value)
                else:
                    raise ValueError(f"XIR syntax error: '{value}' is not a wire.")
                if self.accept("]"):
                    break
                self.expect(",")
        return tuple(wires)

    def application(self) -> Statement:
        name = self.name()
        params = self.params()
        self.expect("|")
        wires = self.wire_list()
        self.expect(";")
        return Statement(name, params, wires)

    def signature(self) -> Tuple[str, List[str], Tuple[Any, ...]]:
        name = self.name()
        params: List[str] = []
        if self.accept("("):
            if not self.accept(")"):
                while True:
                    params.append(self.name())
                    if self.accept(")"):
                        break
                    self.expect(",")
        wires: Tuple[Any, ...] = ()
        if self.peek() == ("op", "["):
            wires = self.wire_list()
        return name, params, wires

    # -- the program
    def program(self) -> Program:
        prog = Program()
        while self.peek()[0] != "eof":
            kind, value = self.peek()
            if kind == "id" and value == "use":
                self.take()
                prog.includes.append(self.take()[1])
                self.expect(";")
            elif kind == "id" and value == "options" and self.peek(1) == ("op", ":"):
                self.take()
                self.take()
                while not self.accept("end"):
                    key = self.name()
                    self.expect(":")
                    prog.options[key] = self.value()
                    self.expect(";")
                self.expect(";")
            elif kind == "id" and value in ("gate", "obs", "out", "func") and self.peek(1)[0] == "id":
                self.take()
                name, params, wires = self.signature()
                if self.accept(";"):
                    prog.add_declaration(Declaration(value, name, params, wires))
                    continue
                self.expect(":")
                if value == "gate":
                    body: List[Statement] = []
                    while not self.accept("end"):
                        body.append(self.application())
                    self.expect(";")
                    if not wires:  # undeclared wires: the ones the body uses, in order of appearance
                        seen: List[Any] = []
                        for stmt in body:
                            seen += [w for w in stmt.wires if w not in seen]
                        wires = tuple(seen)
                    prog.add_declaration(Declaration("gate", name, params, wires))
                    prog.gates[name] = body
                elif value == "obs":
                    terms: List[ObservableStmt] = []
                    while not self.accept("end"):
                        pref = self.value()
                        self.expect(",")
                        factors = []
                        while True:
                            op = self.name()
                            wire = self.wire_list()
                            factors.append((op, wire[0]))
                            if not self.accept("@"):
                                break
                        self.expect(";")
                        terms.append(ObservableStmt(pref, factors))
                    self.expect(";")
                    prog.add_declaration(Declaration("obs", name, params, wires))
                    prog.observables[name] = terms
                else:
                    raise ValueError(f"XIR syntax error: '{value}' definitions are not supported.")
            else:
                prog.statements.append(self.application())
        return prog


def parse_script(text: str, eval_pi: bool = False) -> Program:
    """Parses an XIR script (the subset above)."""
    return _Parser(text, eval_pi).program()
