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


def gather_candidates(records, count, group=None, cap=None):
    """All-gather of variable-length candidate records.

    records : int32 tensor [n, 4] (device tensor for NCCL, host tensor for gloo) whose
              first `count` rows are valid cs_candidate records.
    cap     : None -- two collectives sized by the largest count (one host read of the counts
              in between); an int -- ONE fixed-size exchange of `cap` records per rank plus the
              counts, without any host synchronisation (the caller checks counts <= cap when
              it reads them).
    Returns (tensor [world, max_count or cap, 4], counts tensor [world] on the same device)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = records.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = count if torch.is_tensor(count) else torch.tensor([count], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine.reshape(1).to(torch.int64), group=group)
    mx = int(cap) if cap is not None else max(int(counts.max().item()), 1)
    send = records[:mx]
    if send.shape[0] < mx:  # the local buffer is shorter than the exchanged size
        pad = torch.zeros((mx - send.shape[0], 4), dtype=records.dtype, device=dev)
        send = torch.cat([send, pad])
    out = torch.empty((world * mx, 4), dtype=records.dtype, device=dev)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    return out.view(world, mx, 4), counts


def gather_rows(rows, group=None):
    """Variable-length all-gather of the rows of a 2-D numeric array (same dtype and column
    count on every rank): tensors over NCCL (device) or gloo (host), no pickling.
    Returns the list of per-rank numpy arrays (every rank gets all of them)."""
    import torch
    import torch.distributed as dist
    rows = np.ascontiguousarray(rows)
    if rows.ndim != 2:
        raise ValueError("gather_rows wants a 2-D array")
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [rows]
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
        else torch.device("cpu")
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev), group=group)
    counts = counts.cpu().numpy()
    mx = max(int(counts.max()), 1)
    send = torch.zeros((mx, rows.shape[1]), dtype=torch.from_numpy(rows[:0]).dtype, device=dev)
    if rows.shape[0]:
        send[: rows.shape[0]] = torch.from_numpy(rows).to(dev)
    out = torch.empty((world * mx, rows.shape[1]), dtype=send.dtype, device=dev)
    dist.all_gather_into_tensor(out, send, group=group)
    out = out.cpu().numpy().reshape(world, mx, rows.shape[1])
    return [out[r, : int(counts[r])].copy() for r in range(world)]


def merge_candidates(gathered, counts, row_offsets=None):
    """Gathered records -> one structured numpy array (_lib.CANDIDATE_DTYPE); `row_offsets`
    (one per rank) shifts slab-local coordinates back to chromosome coordinates."""
    g = gathered.cpu().numpy()
    c = counts.cpu().numpy()
    parts = []
    for r in range(g.shape[0]):
        if int(c[r]) > g.shape[1]:
            raise ValueError(f"rank {r} found {int(c[r])} candidates, more than the {g.shape[1]} exchanged")
        rec = np.ascontiguousarray(g[r, : int(c[r])]).view(_lib.CANDIDATE_DTYPE).reshape(-1).copy()
        if row_offsets is not None:
            rec["row"] += int(row_offsets[r])
            rec["col"] += int(row_offsets[r])
        parts.append(rec)
    return np.concatenate(parts) if parts else np.zeros(0, dtype=_lib.CANDIDATE_DTYPE)
