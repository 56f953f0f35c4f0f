#!/usr/bin/env python
"""bench.py — sliced Sycamore-53 amplitude throughput (slices/s) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU contractor

Workload (config.workload): the reference's shipped Sycamore-53 m=12 network + path
(examples/paper_benchmarks/data_files/m12.json) sliced over the 9 indices the reference's own
benchmark uses (examples/paper_benchmarks/CPU/jet_cpu_m12/jet_sliced.cpp:53-54) -> 512 slices.
A "step" = one batch of `slices_per_step` slices per GPU through the hot path
(device-side slice selection -> 167 contraction steps -> FP64 accumulation).  Slices are
partitioned across ranks with no data-path collective; one NCCL reduce of the partial
amplitudes closes the timed region (weak scaling: per-GPU work fixed).

JSON keys follow the driver contract; see DESIGN.md §Measurement for how `roofline`,
`cpu_baseline` and `e2e` are obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DATA_DIR = os.path.join(ROOT, "oracle", "_ref", "data_files")
WORKLOADS = {
    # name: (file, sliced indices, dtype)
    "sycamore53_m12_s9": ("m12.json", "h5 m H10 w y J S G10 P0".split(), "complex64"),
    "sycamore53_m10_s6": ("m10.json", "p7 s7 h4 m1 m2 I2".split(), "complex64"),
    "sycamore53_m10_s10": ("m10.json", "p7 s7 h4 m1 m2 I2 V4 z2 t4 C1".split(), "complex64"),
    # GBS (BASELINE config 5): the reference ships no slice set for these files
    # (examples/paper_benchmarks/CPU/jet_cpu_gbs/jet_gbs_full.cpp contracts unsliced), so the sliced
    # indices are chosen by the greedy slicer (jet_b200/slicing.py): ("auto", number of indices)
    "gbs_fock4_total10_s2": ("gbs_dim2_nc1_lw8_rp5_fock4_total10_0.kraken.json", ("auto", 2), "complex128"),
    "gbs_fock8_total0_s2": ("gbs_dim2_nc1_lw8_rp5_fock8_total0_0.kraken.json", ("auto", 2), "complex128"),
}


def resolve_sliced(workload):
    """The workload's sliced-index list (running the greedy slicer for ("auto", n) entries)."""
    fn, sliced, dt = WORKLOADS[workload]
    if isinstance(sliced, tuple) and sliced[0] == "auto":
        from jet_b200.slicing import find_slices
        js = json.load(open(os.path.join(DATA_DIR, fn)))
        leaf = [t[1] for t in js["tensors"]]
        dims = {}
        for t in js["tensors"]:
            for i, d in zip(t[1], t[2]):
                dims[i] = d
        return find_slices(leaf, dims, [tuple(p) for p in js["path"]], [], extra=sliced[1])
    return list(sliced)
METRIC = "sliced Sycamore-53 amplitude slices/s"  # for the gbs_* workloads: sliced GBS amplitude slices/s
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


# --------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            js = json.load(open(path))
            for key in ("hbm_gbs", "hbm_GBs", "hbm_gb_s"):
                if key in js:
                    return float(js[key]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_network(workload):
    from jet_b200 import NetworkFile
    fn, _, dt = WORKLOADS[workload]
    sliced = resolve_sliced(workload)
    path = os.path.join(DATA_DIR, fn)
    if not os.path.exists(path):
        raise SystemExit(f"{path} missing: run `make -C oracle` where /root/reference exists")
    return NetworkFile.load(path, np.dtype(dt)), sliced, dt, path


# --------------------------------------------------------------------------------------------
# CPU reference leg (test infrastructure: oracle/_ref = the unmodified reference headers)
# --------------------------------------------------------------------------------------------
def cpu_reference_sample(workload, budget_s=20.0):
    """Times the reference's TaskBasedContractor (Taskflow stand-in + OpenBLAS) on the host cores
    over a BOUNDED sample of the workload: one slice is split further by `extra` greedily chosen
    indices; `n` of those sub-slices run concurrently on all host threads.  slices/s is scaled by
    Jet-convention flops: (sub-slice flops / slice flops) * n / seconds."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from jet_b200.slicing import find_slices, replay
    from oracle import ref
    if not ref.available():
        return None
    fn, _, dt = WORKLOADS[workload]
    sliced = resolve_sliced(workload)
    text = open(os.path.join(DATA_DIR, fn)).read()
    js = json.loads(text)
    leaf = [t[1] for t in js["tensors"]]
    dims = {}
    for t in js["tensors"]:
        for i, d in zip(t[1], t[2]):
            dims[i] = d
    path = [tuple(p) for p in js["path"]]
    f_slice, _, _ = replay(leaf, dims, path, sliced)
    cores = os.cpu_count() or 1
    ref.set_blas_threads(1)
    extra = 5 if workload.startswith("sycamore53_m12") else 0
    full = find_slices(leaf, dims, path, sliced, extra=extra) if extra else list(sliced)
    f_sub, _, _ = replay(leaf, dims, path, full)
    sub_per_slice = 2 ** (len(full) - len(sliced))
    # slice id 0 of the workload = sub-slices 0 .. sub_per_slice-1 (the extra indices are the
    # fastest-varying digits of the raveled id)
    t0 = time.time()
    _, sec1, _ = ref.network(text, dt, full, 0, 1, 1, 1)
    waves = max(1, min(4, int(budget_s / (1.5 * max(sec1, 1e-3)))))
    n = min(cores * waves, 512)
    _, sec, _ = ref.network(text, dt, full, 0, 1, cores, n)
    value = (f_sub / f_slice) * n / sec
    return {
        "value": value, "unit": "slices/s", "cores": cores, "kind": "reference",
        "sample": (f"reference TaskBasedContractor (Taskflow stand-in, {cores} threads, OpenBLAS 1 thread/task) on {n} "
                   f"sub-slices of slice 0 ({workload} further sliced over {len(full) - len(sliced)} indices "
                   f"{full[len(sliced):]}: 1/{sub_per_slice} slice each), {sec:.2f} s; scaled by Jet flops "
                   f"{f_sub:.4g}/{f_slice:.4g} per sub-slice"),
        "seconds": sec, "single_subslice_seconds": sec1, "wall_s": time.time() - t0,
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t_all = []
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_reference_sample(args.workload, budget_s=8.0)
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libjetref.so not built"}))
            return
        if i >= args.warmup:
            t_all.append(res)
        if sum(r["wall_s"] for r in t_all) > 150:
            break
    value = float(np.mean([r["value"] for r in t_all]))
    fn, _, dt = WORKLOADS[args.workload]
    sliced = resolve_sliced(args.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": args.gpus,
        "steps": len(t_all), "warmup": args.warmup, "ms_per_step": float(np.mean([r["seconds"] for r in t_all]) * 1e3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32" if dt == "complex64" else "f64",
        "data": "reference data file " + fn,
        "config": {"workload": args.workload, "sliced_indices": sliced, "network": fn},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    line["cpu_baseline"]["value"] = value
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from jet_b200 import ContractionPlan
    from jet_b200.distributed import reduce_amplitude

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    net, sliced, dt, _ = load_network(args.workload)
    plan = ContractionPlan(net, sliced, device=local)
    st = plan.stats
    sps = args.slices_per_step
    total = plan.num_slices
    stream = torch.cuda.ExternalStream(plan.stream(), device=torch.device("cuda", local))

    def slice_ids(step_no):
        base = ((step_no * world + rank) * sps) % total
        return [(base + j) % total for j in range(sps)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # pinned host copies of the leaves: the host->device leg of the end-to-end run
    leaves = [torch.from_numpy(np.ascontiguousarray(arr)).pin_memory() for _, arr in net.tensors]
    leaf_ptrs = [t.data_ptr() for t in leaves]
    h2d_bytes = int(sum(t.numel() * t.element_size() for t in leaves))

    # ---- device-resident arm -------------------------------------------------------------
    plan.reset()
    for w in range(args.warmup):
        plan.run_list(slice_ids(w))
    plan.sync()
    plan.reset()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(args.steps):
        plan.run_list(slice_ids(args.warmup + k))
    if world > 1:
        # the one exchange step: NCCL reduce of the FP64 partial amplitudes to rank 0
        reduce_amplitude(plan.result(), dst=0, device=torch.device("cuda", local))
    cur = torch.cuda.current_stream()
    cur.wait_stream(stream)
    e1.record(cur)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    slices_done = args.steps * sps * world
    value = slices_done / (max_ms * 1e-3)
    amp = plan.result().reshape(-1)[0]

    # ---- end-to-end arm: pinned host leaves -> device, run, result -> host, every step ------
    plan.reset()
    for w in range(min(args.warmup, 3)):
        plan.upload_ptrs(leaf_ptrs)
        plan.reset()
        plan.run_list(slice_ids(w))
        plan.result()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        plan.upload_ptrs(leaf_ptrs)  # H2D of this step's inputs (all leaves)
        plan.reset()                 # shared (slice-independent) subtrees are recomputed
        plan.run_list(slice_ids(args.warmup + k))
        r = plan.result()            # D2H of the step's accumulated amplitude (syncs)
    if world > 1:
        reduce_amplitude(r, dst=0, device=torch.device("cuda", local))
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = slices_done / (float(t.item()) * 1e-3)
    d2h_bytes = 16 * plan.result_elems

    # ---- roofline of the dominant kernel (per-step CUDA-event profile, outside the timed region)
    peak, peak_src = measured_peaks()
    line = None
    if rank == 0:
        prof = plan.profile(slice_ids(0)[0], 3)
        steps = plan.steps()
        stream_ms = sum(float(prof[i]) for i, s in enumerate(steps) if not s.shared and s.kernel == 0)
        stream_bytes = sum(s.bytes for s in steps if not s.shared and s.kernel == 0)
        stream_launches = sum(1 for s in steps if not s.shared and s.kernel == 0)
        all_ms = float(prof.sum())
        # whole-step figure: algorithmic bytes of one slice / measured time of one slice in the
        # timed region (includes launch gaps and the non-dominant kernels: conservative)
        ms_per_slice = max_ms / (args.steps * sps)
        whole_slice = st.bytes_per_slice / (ms_per_slice * 1e-3) / 1e9
        achieved = stream_bytes / (stream_ms * 1e-3) / 1e9 if stream_ms > 0 else 0.0
        traffic, traffic_detail = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic_detail = json.load(open(tpath)).get(args.workload)
                ratio = traffic_detail["dram_bytes_per_launch"] / traffic_detail["algorithmic_bytes_same_launch"]
                traffic = ratio * stream_bytes / max(stream_launches, 1)
            except Exception:
                traffic = None
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": peak_src, "kernel": "StreamContractKernel",
            "definition": "sum of algorithmic bytes sizeof(T)*(MK+KN+MN) over the StreamContractKernel launches of one "
                          "slice / sum of their durations, each launch timed with CUDA events on the plan's stream "
                          "(jb_plan_profile, 3 repetitions, run right after the timed region)",
            "launches_per_slice": stream_launches, "share_of_slice_time": stream_ms / all_ms if all_ms else None,
            "bytes_per_launch_avg": stream_bytes / max(stream_launches, 1),
            "ms_per_launch_avg": stream_ms / max(stream_launches, 1),
            "traffic_detail": traffic_detail,
            "whole_slice": {"GBs": whole_slice, "frac": whole_slice / peak,
                            "definition": "algorithmic bytes of one slice / device time per slice inside the timed region "
                                          "(includes launch gaps and the non-dominant kernels)"},
        }
        cpu = None if args.no_cpu else cpu_reference_sample(args.workload)
        fn = WORKLOADS[args.workload][0]
        line = {
            "metric": METRIC, "value": value, "unit": "slices/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if dt == "complex64" else "f64",
            "data": f"reference data file {fn} (network + path), slice ids synthetic",
            "config": {"workload": args.workload, "network": fn, "sliced_indices": sliced, "num_slices": total,
                       "slices_per_step_per_gpu": sps, "steps_per_slice": int(st.steps_total - st.steps_shared),
                       "l2": "per-slice working set (%.1f GB algorithmic) exceeds L2; no flush needed" % (st.bytes_per_slice / 1e9),
                       "parallelism": f"slices partitioned over {world} GPU(s), one NCCL reduce"},
            "tflops": st.flops_per_slice * slices_done / (max_ms * 1e-3) / 1e12,
            "amplitude_partial": [float(amp.real), float(amp.imag)],
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "slices/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": int(st.launches_per_slice) * args.steps * sps,
            "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    plan.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sycamore53_m12_s9", choices=sorted(WORKLOADS))
    ap.add_argument("--slices-per-step", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
