"""Multi-GPU host logic: slices are the data-parallel unit.  One process per GPU
(torchrun); rank r contracts a contiguous block of slice ids on its own device with no data-path
collective, and ONE reduce of the double-precision partial amplitudes closes the run — NCCL over
NVLink on GPUs, gloo in the CPU tests.  The reference has no inter-device path at all
(reference include/jet/TaskBasedContractor.hpp:258-280 reduces inside one process)."""
from __future__ import annotations
# Two reduce paths: `reduce_on_device` (the product path on GPUs: jb_multi_reduce = ncclReduce of the FP64
# on-device total, enqueued on the plan's stream, no host staging; the communicator is jet_b200.Communicator,
# its id handed around by torch.distributed) and `reduce_amplitude` (host arrays through torch.distributed:
# what the CPU gloo tests exercise, and a cross-check of the first on GPUs).

from typing import Optional, Tuple

import numpy as np


def slice_range(num_slices: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous partition of [0, num_slices) over `world` ranks; the first (num_slices % world)
    ranks get one extra slice.  Returns (first, count)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("invalid rank/world")
    base, extra = divmod(num_slices, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def reduce_amplitude(partial: np.ndarray, dst: int = 0, device=None) -> Optional[np.ndarray]:
    """Sum the per-rank partial results (complex128 array) onto rank `dst` with one collective.
    Returns the total on `dst`, None elsewhere.  Without an initialised process group (single
    process) the input is returned unchanged."""
    import torch
    import torch.distributed as dist

    partial = np.ascontiguousarray(partial, dtype=np.complex128)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return partial
    t = torch.from_numpy(partial.reshape(-1).view(np.float64).copy())
    if device is not None:
        t = t.to(device)
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    if dist.get_rank() != dst:
        return None
    return t.cpu().numpy().view(np.complex128).reshape(partial.shape)


def make_communicator(device: int):
    """A jet_b200.Communicator (NCCL, bound inside libjetb200.so) spanning the initialised torch.distributed
    group: rank 0's unique id travels as a broadcast object."""
    import torch.distributed as dist

    from .plan import Communicator

    def exchange(ident):
        box = [ident]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    return Communicator(dist.get_world_size(), dist.get_rank(), device, exchange)


def reduce_on_device(plan, comm, root: int = 0) -> Optional[np.ndarray]:
    """Finish a sliced run: the plan set's on-device FP64 total is reduced over the ranks by one ncclReduce on
    the plan's stream; only `root` copies the 16 x result_elems bytes back.  Returns the total on root."""
    plan.reduce(comm, root)
    if comm.rank != root:
        plan.sync()
        return None
    return plan.result()
