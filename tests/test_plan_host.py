"""Host-side planning without a GPU (JB_PLAN_DRY_RUN): the plan of the reference's m10 network — shared / per-slice split,
fused chains, slice batching decision, arena layout, launch units — is computed on a CPU-only machine; everything that
would touch the device fails loudly (there is no CPU execution path)."""
import os

import numpy as np
import pytest

from jet_b200 import ContractionPlan, NetworkFile
from jet_b200._lib import JetB200Error


def test_dry_run_plan_of_m10(data_dir):
    net = NetworkFile.load(os.path.join(data_dir, "m10.json"), np.complex64)
    sliced = "p7 s7 h4 m1 m2 I2 V4 z2 t4 C1".split()
    with ContractionPlan(net, sliced, dry_run=True) as plan:
        st = plan.stats
        assert plan.num_slices == 1024 and st.steps_total == 321
        # the reference's name de-duplication (TaskBasedContractor.hpp:216-222) as a plan-time split
        assert 0 < st.steps_shared < st.steps_total
        assert st.chains > 0 and st.steps_chained > 0 and st.fused_bytes_per_slice < st.bytes_per_slice
        assert st.batch == 64  # per-slice tensors are small: 64 slices per launch
        units = plan.ops()
        assert len(units) >= st.chains and sum(u.n_steps for u in units) == st.steps_total - st.steps_shared
        assert all(u.register_steps <= u.n_steps for u in units)
        steps = plan.steps()
        assert len(steps) == 321 and steps[-1].shared == 0
        with pytest.raises(JetB200Error, match="JB_PLAN_DRY_RUN"):
            plan.reset()
        with pytest.raises(JetB200Error, match="JB_PLAN_DRY_RUN"):
            plan.run(0, 1)
    with ContractionPlan(net, sliced, dry_run=True, batch=1) as plan:
        assert plan.stats.batch == 1
    # Jet-convention flops of the plan = PathInfo's (2*M*N*K over all steps of the sliced network)
    with ContractionPlan(net, sliced[:6], dry_run=True) as plan:
        import json
        gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "amplitudes.json")))
        assert plan.stats.jet_flops_per_slice == gold["m10_s6_slice0_complex64"]["jet_flops"]


def test_dry_run_plan_reports_errors_like_the_reference(data_dir):
    net = NetworkFile.load(os.path.join(data_dir, "m10.json"), np.complex64)
    with pytest.raises(ValueError, match="Sliced index does not exist."):
        ContractionPlan(net, ["no_such_index"], dry_run=True)
    bad = NetworkFile(net.tensors, [(0, 100000)] + net.path[1:])
    with pytest.raises(JetB200Error, match="Node ID 2 in contraction pair is invalid."):
        ContractionPlan(bad, [], dry_run=True)
