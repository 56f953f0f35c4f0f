#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_m20_synth.py::test_m20_per_step_normwise_vs_reference > gpurun_out/r2h_pytest.log 2>&1; tail -12 gpurun_out/r2h_pytest.log
for wl in sycamore53_m10_s10 sycamore53_m10_s6 gbs_fock4_total10_s2; do
  for lanes in 1 2 4; do
    echo "== $wl lanes $lanes"
    timeout 300 python bench.py --workload $wl --no-cpu --no-others --steps 4 --warmup 2 --slices-per-step 256 --lanes $lanes --strong-slices 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['config']['slices_per_launch'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline'].get('whole_slice',{}).get('frac'))"
  done
  echo "== $wl lanes 4 batch off"
  JB_PLAN_BATCH=1 timeout 300 python bench.py --workload $wl --no-cpu --no-others --steps 4 --warmup 2 --slices-per-step 256 --lanes 4 --strong-slices 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['config']['slices_per_launch'])"
done
