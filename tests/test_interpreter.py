"""The XIR route of the Python front end (jet_b200/xir_lite.py + interpreter.py) against the reference's interpreter
tests (python/tests/test_interpreter.py, restated: same scripts, same expected values and error messages).  Parsing,
the manifest and every validation error run on the CPU (they are raised before anything is contracted); the programs
that produce values run on the GPU (each output statement = one plan)."""
from inspect import cleandoc
from math import sqrt

import numpy as np
import pytest

from jet_b200 import interpreter as ji
from jet_b200.xir_lite import parse_script


# ---------------------------------------------------------------- parsing
def test_statement_text_round_trip():
    prog = parse_script("""
        use <xc/jet>;
        options: dimension: 3; end;
        gate RX3 (a, b, c) [0, 1, 2]:
            RY3(a: b, b: c, c: a) | [0, 1, 2];
        end;
        obs XY[0, 1]: 1, X[0]; 2.5, Y[1] @ Z[0]; end;
        RX3(a: 0, b: 3.25, c: pi) | [0, 1, 2];
        amplitude(state: [0, -1]) | [0, 1];
        amplitude(state) | [0];
        Negate(0) | [4];
        probabilities | [0, 1, 2];
    """)
    assert prog.options == {"dimension": 3}
    assert [str(s) for s in prog.statements] == [
        "RX3(a: 0, b: 3.25, c: PI) | [0, 1, 2]", "amplitude(state: [0, -1]) | [0, 1]", "amplitude(state) | [0]",
        "Negate(0) | [4]", "probabilities | [0, 1, 2]"]
    assert prog.wires == [0, 1, 2, 4]
    assert [str(s) for s in prog.gates["RX3"]] == ["RY3(a: b, b: c, c: a) | [0, 1, 2]"]
    assert [str(s) for s in prog.observables["XY"]] == ["1, X[0]", "2.5, Y[1] @ Z[0]"]
    assert prog.search("gate", "params", "RX3") == ["a", "b", "c"] and prog.search("obs", "wires", "XY") == (0, 1)
    assert parse_script("RY(pi/2) | [0];", eval_pi=True).statements[0].params == [pytest.approx(np.pi / 2)]
    assert str(parse_script("RY(pi/2) | [0];").statements[0]) == "RY(PI/2) | [0]"
    with pytest.raises(ValueError, match="XIR syntax error"):
        parse_script("H | 0;")


def test_manifest_lists_every_gate_and_output():
    # python/tests/test_interpreter.py:96-191 (first and last entries, and the parameter / wire forms)
    text = ji.get_xir_manifest().serialize(minimize=True)
    assert text.startswith("gate BS(theta, phi)[0, 1]; gate Beamsplitter(theta, phi)[0, 1]; gate CNOT[0, 1]; "
                           "gate CPhaseShift(phi)[0, 1]; gate CRX(theta)[0, 1];")
    assert text.endswith("gate x[0]; gate y[0]; gate z[0]; out Amplitude; out Expval; out Probabilities; "
                         "out amplitude; out expval; out probabilities;")
    for piece in ("gate CRot(phi, theta, omega)[0, 1];", "gate CSWAP[0, 1, 2];", "gate D(r, phi)[0];",
                  "gate TwoModeSqueezing(r, theta)[0, 1];", "gate U3(theta, phi, lam)[0];", "gate toffoli[0, 1, 2];"):
        assert piece in text
    assert text.count("gate ") == 78 and text.count("out ") == 6


# ---------------------------------------------------------------- programs that stop before anything is contracted
@pytest.mark.parametrize("script", ["", "use <xc/jet>;", "H | [0];"])
def test_programs_without_output_statements(script):
    assert ji.run_xir_program(parse_script(script)) == []


@pytest.mark.parametrize("script, match", [
    ("options: dimension: [2]; end;", r"Option 'dimension' must be an integer\."),
    ("options: dimension: 1; end;", r"Option 'dimension' must be greater than one\."),
    ("options: dimension: 3; end; X | [0];",
     r"Statement 'X \| \[0\]' applies a gate with a dimension \(2\) that differs from the dimension of the circuit \(3\)\."),
    ("X | [0]; amplitude | [0];", r"Statement 'amplitude \| \[0\]' is missing a 'state' parameter\."),
    ("X | [0]; amplitude(state) | [0];", r"Statement 'amplitude\(state\) \| \[0\]' is missing a 'state' parameter\."),
    ("X | [0]; amplitude(state: [0, -1]) | [0, 1];",
     r"Statement 'amplitude\(state: \[0, -1\]\) \| \[0, 1\]' has a 'state' parameter with at least one entry that falls "
     r"outside the range \[0, 2\)\."),
    ("X | [0]; amplitude(state: [0, 2]) | [0, 1];",
     r"Statement 'amplitude\(state: \[0, 2\]\) \| \[0, 1\]' has a 'state' parameter with at least one entry that falls "
     r"outside the range \[0, 2\)\."),
    ("X | [0]; amplitude(state: [0, 0]) | [0];",
     r"Statement 'amplitude\(state: \[0, 0\]\) \| \[0\]' has a 'state' parameter with 2 \(!= 1\) entries\."),
    ("X | [0]; amplitude(state: [0]) | [0, 1];",
     r"Statement 'amplitude\(state: \[0\]\) \| \[0, 1\]' has a 'state' parameter with 1 \(!= 2\) entries\."),
    ("CNOT | [0, 1]; amplitude(state: [0, 1]) | [0];",
     r"Statement 'amplitude\(state: \[0, 1\]\) \| \[0\]' must be applied to \[0 \.\. 1\]\."),
    ("CNOT | [0, 1]; amplitude(state: [0, 1]) | [1, 0];",
     r"Statement 'amplitude\(state: \[0, 1\]\) \| \[1, 0\]' must be applied to \[0 \.\. 1\]\."),
    ("CNOT | [0, 1]; probabilities | [0];", r"Statement 'probabilities \| \[0\]' must be applied to \[0 \.\. 1\]\."),
    ("gate Circle[0]: Circle | [0]; end; Circle | [0];", r"Gate 'Circle' has a circular dependency\."),
    ("gate Day[0]: Dawn | [0]; end; gate Dawn[0]: Dusk | [0]; end; gate Dusk[0]: Dawn | [0]; end; Day | [0];",
     r"Gate 'Dawn' has a circular dependency\."),
    ("gate Incomplete[0]: Missing | [0]; end; Incomplete | [0];",
     r"Statement 'Missing \| \[0\]' applies a gate which has not been defined\."),
    ("gate Negate [0]: X | [0]; end; Negate(0) | [0];",
     r"Statement 'Negate\(0\) \| \[0\]' has the wrong number of parameters\."),
    ("gate Spin(theta) [0]: Rot(theta, theta, theta) | [0]; end; Spin(phi: pi) | [0];",
     r"Statement 'Spin\(phi: PI\) \| \[0\]' has an invalid set of parameters\."),
    ("gate Permute [0, 1]: SWAP | [0, 1]; end; Permute | [0];",
     r"Statement 'Permute \| \[0\]' has the wrong number of wires\."),
    ("expval | [0];", r"Statement 'expval \| \[0\]' is missing an 'observable' parameter\."),
    ("expval(observable: dne) | [0];",
     r"Statement 'expval\(observable: dne\) \| \[0\]' has an 'observable' parameter which references an undefined observable\."),
    ("obs box[0]; expval(observable: box) | [0];",
     r"Statement 'expval\(observable: box\) \| \[0\]' has an 'observable' parameter which references an undefined observable\."),
    ("obs up(scale)[0]: scale, Z[0]; end; expval(observable: up) | [0];",
     r"Statement 'expval\(observable: up\) \| \[0\]' has an 'observable' parameter which references a parameterized observable\."),
    ("obs obs[0]: 1, Z[0]; end; X | [0]; X | [1]; expval(observable: obs) | [0, 1];",
     r"Statement 'expval\(observable: obs\) \| \[0, 1\]' has an 'observable' parameter which applies the wrong number of wires\."),
    ("obs natural[0]: one, Z[0]; end; expval(observable: natural) | [0];",
     r"Observable statement 'one, Z\[0\]' has a prefactor \(one\) which cannot be converted to a floating-point number\."),
    ("halt | [0];", r"Statement 'halt \| \[0\]' is not supported\."),
])
def test_invalid_programs(script, match):
    with pytest.raises(ValueError, match=match):
        ji.run_xir_program(parse_script(script))


def test_unsupported_option_warns():
    with pytest.warns(UserWarning, match=r"Option 'VSync' is not supported and will be ignored\."):
        ji.run_xir_program(parse_script("options: dimension: 3; VSync: off; end;"))


# ---------------------------------------------------------------- programs with values (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("script, want", [
    # python/tests/test_interpreter.py:203-262, 309-357
    ("""use <xc/jet>; options: dimension: 2; end;
        H | [0]; S | [0]; Displacement(3, 1) | [0];
        amplitude(state: [0]) | [0]; amplitude(state: [1]) | [0];""",
     [-0.011974639958 - 0.012732623852j, 0.012732623852 - 0.043012087532j]),
    ("""use <xc/jet>; options: dimension: 3; end;
        Squeezing(1, 2) | [0]; Squeezing(2, 1) | [1];
        amplitude(state: [0, 0]) | [0, 1]; amplitude(state: [0, 1]) | [0, 1]; amplitude(state: [0, 2]) | [0, 1];
        amplitude(state: [1, 0]) | [0, 1]; amplitude(state: [1, 1]) | [0, 1]; amplitude(state: [1, 2]) | [0, 1];
        amplitude(state: [2, 0]) | [0, 1]; amplitude(state: [2, 1]) | [0, 1]; amplitude(state: [2, 2]) | [0, 1];""",
     [0.415035263978, 0, -0.152860853701 - 0.238066674351j, 0, 0, 0, 0.093012260922 - 0.203235497887j, 0,
      -0.150834249845 + 0.021500900893j]),
    ("X | [0]; amplitude(state: [0]) | [0]; amplitude(state: [1]) | [0];", [0, 1]),
    ("""H | [0]; CNOT | [0, 1];
        amplitude(state: [0, 0]) | [0, 1]; amplitude(state: [0, 1]) | [0, 1];
        amplitude(state: [1, 0]) | [0, 1]; amplitude(state: [1, 1]) | [0, 1];""", [1 / sqrt(2), 0, 0, 1 / sqrt(2)]),
    ("""TwoModeSqueezing(3, 1) | [0, 1];
        amplitude(state: [0, 0]) | [0, 1]; amplitude(state: [0, 1]) | [0, 1];
        amplitude(state: [1, 0]) | [0, 1]; amplitude(state: [1, 1]) | [0, 1];""",
     [0.0993279274194332, 0, 0, 0.053401711152745175 + 0.08316823745907517j]),
    # gate definitions, :504-572
    ("gate H2: H | [0]; H | [1]; end; X | [0]; amplitude(state: [0]) | [0]; amplitude(state: [1]) | [0];", [0, 1]),
    ("""gate H2 [0, 1]: H | [0]; H | [1]; end; H2 | [0, 1];
        amplitude(state: [0, 0]) | [0, 1]; amplitude(state: [0, 1]) | [0, 1];
        amplitude(state: [1, 0]) | [0, 1]; amplitude(state: [1, 1]) | [0, 1];""", [0.5, 0.5, 0.5, 0.5]),
    ("""gate Flip[coin]: X | [coin]; end;
        gate Stay[coin]: Flip | [coin]; Flip | [coin]; end;
        gate FlipStay[0, 1]: Flip | [0]; Stay | [1]; end;
        FlipStay | [0, 1];
        amplitude(state: [0, 0]) | [0, 1]; amplitude(state: [0, 1]) | [0, 1];
        amplitude(state: [1, 0]) | [0, 1]; amplitude(state: [1, 1]) | [0, 1];""", [0, 0, 1, 0]),
    ("""gate RY3 (a, b, c) [0, 1, 2]: RY(a) | [0]; RY(b) | [1]; RY(c) | [2]; end;
        gate RX3 (a, b, c) [0, 1, 2]: RY3(a: b, b: c, c: a) | [0, 1, 2]; end;
        RX3(a: 0, b: 3.141592653589793, c: 3.141592653589793) | [0, 1, 2];
        amplitude(state: [0, 0, 0]) | [0, 1, 2]; amplitude(state: [0, 0, 1]) | [0, 1, 2];
        amplitude(state: [0, 1, 0]) | [0, 1, 2]; amplitude(state: [0, 1, 1]) | [0, 1, 2];
        amplitude(state: [1, 0, 0]) | [0, 1, 2]; amplitude(state: [1, 0, 1]) | [0, 1, 2];
        amplitude(state: [1, 1, 0]) | [0, 1, 2]; amplitude(state: [1, 1, 1]) | [0, 1, 2];""", [0, 0, 0, 0, 0, 0, 1, 0]),
    # expected values, :690-737
    ("obs Z [0]: 1, Z[0]; end; expval(observable: Z) | [0];", [1]),
    ("obs Z3 [wire]: 3, Z[wire]; end; X | [0]; expval(observable: Z3) | [0];", [-3]),
    ("""obs XY[0, 1]: 1, X[0]; 1, Y[1]; end; obs YX[0, 1]: 1, Y[0]; 1, X[1]; end;
        RY(pi/2) | [0]; RX(pi/4) | [1];
        expval(observable: XY) | [0, 1]; expval(observable: YX) | [0, 1];""", [-1 / sqrt(2), 0]),
])
def test_programs_with_values(script, want):
    got = ji.run_xir_program(parse_script(cleandoc(script), eval_pi=True))
    assert len(got) == len(want)
    assert [complex(g) for g in got] == pytest.approx([complex(w) for w in want], abs=1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("script, want", [
    # :425-480
    ("probabilities | [0];", [1, 0]),
    ("probabilities | [0, 1, 2];", [1, 0, 0, 0, 0, 0, 0, 0]),
    ("X | [0]; probabilities | [0];", [0, 1]),
    ("X | [1]; probabilities | [0, 1];", [0, 1, 0, 0]),
    ("H | [0]; CNOT | [0, 1]; probabilities | [0, 1];", [0.5, 0, 0, 0.5]),
    ("H | [0]; Y | [0]; probabilities | [0];", [0.5, 0.5]),
])
def test_programs_with_probabilities(script, want):
    got = ji.run_xir_script(script)
    assert len(got) == 1 and np.asarray(got[0]).real == pytest.approx(want, abs=1e-12)
