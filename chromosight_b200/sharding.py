"""Multi-GPU plumbing of the hot path: one process per GPU, no
data-path collective.

The reference hands every intra- or inter-chromosomal sub-matrix to a worker
of a multiprocessing pool and concatenates the per-sub-matrix pattern tables in
the parent (cli:738-804).  Here a rank takes the place of a pool worker:

* `partition_units`  -- which sub-matrices a rank processes (longest processing
  time first, by window count);
* `row_slabs`        -- split ONE large chromosome into contiguous row slabs with
  the kernel halo, for the 1/2/4/8-GPU runs on a single map;
* `gather_candidates` -- the one collective of the path: every rank's candidate
  records (row, col, score, log10 p; 16 B each) to every rank, as two
  all-gathers (counts, then records padded to the largest count).  Works on
  NCCL (device tensors) and gloo (host tensors, used by the CPU tests).
"""
import numpy as np

from . import _lib


def partition_units(costs, world_size):
    """Greedy LPT: units sorted by decreasing cost, each to the least loaded rank.

    costs : sequence of numbers (windows of each sub-matrix: (D+1)*n intra, ms*ns inter).
    Returns a list of `world_size` lists of unit indices (each sorted ascending)."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world_size)
    out = [[] for _ in range(world_size)]
    for u in order:
        r = int(np.argmin(load))
        out[r].append(int(u))
        load[r] += costs[u]
    return [sorted(x) for x in out]


def row_slabs(n_rows, world_size, halo):
    """Contiguous row slabs of one chromosome: rank g scores rows [r0, r1) and needs
    input rows [max(r0 - halo, 0), min(r1 + halo, n_rows)).
    Returns a list of (r0, r1, in0, in1)."""
    bounds = np.linspace(0, n_rows, world_size + 1).round().astype(int)
    out = []
    for g in range(world_size):
        r0, r1 = int(bounds[g]), int(bounds[g + 1])
        out.append((r0, r1, max(r0 - halo, 0), min(r1 + halo, n_rows)))
    return out


def gather_candidates(records, count, group=None):
    """All-gather of variable-length candidate records.

    records : int32 tensor [cap, 4] (device tensor for NCCL, host tensor for gloo) whose
              first `count` rows are valid cs_candidate records.
    Returns (tensor [world, max_count, 4], counts tensor [world] on the same device)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = records.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([count], dtype=torch.int64, device=dev), group=group)
    mx = max(int(counts.max().item()), 1)
    send = records[:mx]
    if send.shape[0] < mx:  # the local buffer is shorter than another rank's count
        pad = torch.zeros((mx - send.shape[0], 4), dtype=records.dtype, device=dev)
        send = torch.cat([send, pad])
    out = torch.empty((world * mx, 4), dtype=records.dtype, device=dev)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    return out.view(world, mx, 4), counts


def merge_candidates(gathered, counts, row_offsets=None):
    """Gathered records -> one structured numpy array (_lib.CANDIDATE_DTYPE); `row_offsets`
    (one per rank) shifts slab-local coordinates back to chromosome coordinates."""
    g = gathered.cpu().numpy()
    c = counts.cpu().numpy()
    parts = []
    for r in range(g.shape[0]):
        rec = np.ascontiguousarray(g[r, : int(c[r])]).view(_lib.CANDIDATE_DTYPE).reshape(-1).copy()
        if row_offsets is not None:
            rec["row"] += int(row_offsets[r])
            rec["col"] += int(row_offsets[r])
        parts.append(rec)
    return np.concatenate(parts) if parts else np.zeros(0, dtype=_lib.CANDIDATE_DTYPE)
