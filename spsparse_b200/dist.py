"""Multi-GPU plumbing for the row-partitioned multiply: A is split by contiguous
row ranges, B is sharded the same way (by its row = the inner index), consolidated locally, and then
replicated once on every rank.  Concatenating the ranks' consolidated blocks of C in rank order IS the
reference's row-major order, so nothing is exchanged after the multiply.

Device-agnostic on purpose: bench.py drives it with CUDA tensors over NCCL (NVLink), the CPU tests
with gloo.  The reference itself has no counterpart (it is single-process)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def row_range(m: int, rank: int, world: int) -> tuple[int, int]:
    """Rows [r0, r1) owned by `rank`."""
    return rank * m // world, (rank + 1) * m // world


def replicate_start(local: list[torch.Tensor], rank: int, world: int):
    """Starts replicating per-rank shards (one tensor per array of the COO: idx0, idx1, val; all of the
    same length on a rank, lengths may differ between ranks).  Each shard is broadcast straight into its
    slice of the full array -- no padding, no compaction copy.  Returns (full_arrays, pending_works, sizes)."""
    dev = local[0].device
    n_local = int(local[0].shape[0])
    sizes_t = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes_t, torch.tensor([n_local], dtype=torch.int64, device=dev))
    sizes = [int(x) for x in sizes_t.tolist()]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    total = int(offs[-1])
    fulls, works = [], []
    for t in local:
        full = torch.empty(total, dtype=t.dtype, device=dev)
        full[offs[rank]:offs[rank + 1]] = t
        for g in range(world):
            if sizes[g]:
                works.append(dist.broadcast(full[offs[g]:offs[g + 1]], src=g, async_op=True))
        fulls.append(full)
    return fulls, works, sizes


def replicate_wait(works) -> None:
    for w in works:
        w.wait()
