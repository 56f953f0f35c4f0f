"""Generate data/<stem>.golden.json (default stem: sycamore53_m20): slices of the synthetic m=20 network contracted by the
UNMODIFIED reference (oracle/_ref: TaskBasedContractor + deletion tasks) in complex64 and complex128.
Run in the container that has /root/reference (several minutes per slice and dtype):
    python tools/make_m20_golden.py [--stem sycamore53_m20] [slice ids ...]
The complex128 value is the reference's own higher-precision result for the same slice: a single
slice amplitude is a sum with heavy cancellation, so two correct complex64 engines differ by more
than 1e-5 on it (the reference's complex64 result is 2e-4 away from its complex128 one on slice 0);
tests compare against both (SURVEY 8c caveat ii)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

DATA = os.path.join(ROOT, "data")
argv = sys.argv[1:]
stem = "sycamore53_m20"
if argv and argv[0] == "--stem":
    stem, argv = argv[1], argv[2:]
ids = [int(a) for a in argv] or [0, 12345678]
meta = json.load(open(os.path.join(DATA, stem + ".meta.json")))
text = open(os.path.join(DATA, stem + ".json")).read()
out_path = os.path.join(DATA, stem + ".golden.json")
gold = json.load(open(out_path)) if os.path.exists(out_path) else {}
ref.set_blas_threads(1)
for v in ids:
    e = gold.setdefault(str(v), {})
    for dt, key in (("complex64", ""), ("complex128", "_c128")):
        if ("re" + key) in e:
            continue
        r, sec, fl = ref.network(text, dt, meta["sliced_indices"], v, mode=2, threads=8, num_slices=1)
        e["re" + key], e["im" + key] = float(r[0].real), float(r[0].imag)
        e["jet_flops"], e["ref_seconds_here" + key] = fl, sec
        print(v, dt, r[0], sec, flush=True)
        json.dump(gold, open(out_path, "w"), indent=1)
