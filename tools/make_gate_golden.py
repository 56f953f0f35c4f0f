"""Golden matrices of the reference's qubit gates (python/jet/gate.py) for tests/test_frontend.py.

The reference package cannot be imported as shipped in this container (its compiled bindings, `thewalrus` and `xir`
are absent), but `gate.py` only needs `thewalrus` for the four Fock gates and `.factory` for `tensor()`: both are
replaced by empty stand-ins here and the module is loaded from where it lies.  Every registered qubit gate is
instantiated with seeded parameters and its `_data()` matrix written to tests/golden/gates.npz together with the
parameters.  (The Fock gates are pinned by the known answers of the reference's own tests instead,
python/tests/test_gate.py:258-381, restated in tests/test_frontend.py.)

    python tools/make_gate_golden.py        # needs /root/reference
"""
import importlib.util
import inspect
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/python/jet/gate.py"


def load_reference_gates():
    walrus = types.ModuleType("thewalrus")
    grads = types.ModuleType("thewalrus.fock_gradients")
    for name in ("beamsplitter", "displacement", "squeezing", "two_mode_squeezing"):
        setattr(grads, name, None)
    walrus.fock_gradients = grads
    sys.modules.setdefault("thewalrus", walrus)
    sys.modules.setdefault("thewalrus.fock_gradients", grads)
    pkg = types.ModuleType("refjet")
    pkg.__path__ = []
    factory = types.ModuleType("refjet.factory")
    factory.Tensor = None
    factory.TensorType = None
    sys.modules["refjet"] = pkg
    sys.modules["refjet.factory"] = factory
    spec = importlib.util.spec_from_file_location("refjet.gate", REF)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refjet.gate"] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_gates()
    rng = np.random.default_rng(2024)
    out = {}
    classes = sorted({cls for cls in ref.GateFactory.registry.values() if issubclass(cls, ref.QubitGate)},
                     key=lambda c: c.__name__)
    for cls in classes:
        n_params = len(inspect.signature(cls.__init__).parameters) - 1
        for rep in range(3 if n_params else 1):
            params = rng.uniform(-np.pi, np.pi, n_params)
            gate = cls(*params)
            key = f"{cls.__name__}/{rep}"
            out[key + "/params"] = params
            out[key + "/matrix"] = np.asarray(gate._data(), dtype=np.complex128)
            out[key + "/adjoint"] = np.asarray(ref.Adjoint(gate)._data(), dtype=np.complex128)
            out[key + "/num_wires"] = np.array(gate.num_wires)
    names = {name: cls.__name__ for name, cls in ref.GateFactory.registry.items()}
    out["registry/names"] = np.array(sorted(names))
    out["registry/classes"] = np.array([names[k] for k in sorted(names)])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gates.npz"), **out)
    print(len(classes), "qubit gate classes,", len(names), "registered names ->", "tests/golden/gates.npz")


if __name__ == "__main__":
    main()
