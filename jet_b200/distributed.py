"""Multi-GPU host logic: slices are the data-parallel unit.  One process per GPU
(torchrun); rank r contracts a contiguous block of slice ids on its own device with no data-path
collective, and ONE reduce of the double-precision partial amplitudes closes the run — NCCL over
NVLink on GPUs, gloo in the CPU tests.  The reference has no inter-device path at all
(reference include/jet/TaskBasedContractor.hpp:258-280 reduces inside one process)."""
from __future__ import annotations

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
