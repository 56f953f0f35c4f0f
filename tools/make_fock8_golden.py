"""Reference result for the GBS fock-8 benchmark file (complex128, unsliced), computed by the UNMODIFIED reference
(oracle/_ref: TaskBasedContractor + deletion tasks).  Run where /root/reference exists (minutes, ~10 GB):
    python tools/make_fock8_golden.py
Writes tests/golden/amplitudes_fock8.json."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

fn = "gbs_dim2_nc1_lw8_rp5_fock8_total0_0.kraken.json"
text = open(os.path.join(ROOT, "data", "_ref", fn)).read()
ref.set_blas_threads(8)
r, sec, fl = ref.network(text, "complex128", [], 0, mode=2, threads=1, num_slices=1)
out = {"gbs_fock8_total0_complex128": {"re": float(r[0].real), "im": float(r[0].imag), "jet_flops": fl,
                                       "ref_seconds_here": sec, "file": fn}}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "amplitudes_fock8.json"), "w"), indent=1)
print(out)
