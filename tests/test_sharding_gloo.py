"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chromosight_b200 import _lib, sharding


def test_partition_units_balances_and_covers():
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 1000, size=276)          # 23 intra + 253 inter sub-matrices
    parts = sharding.partition_units(costs, 8)
    flat = sorted(i for p in parts for i in p)
    assert flat == list(range(276))
    loads = np.array([costs[p].sum() for p in parts])
    assert loads.max() - loads.min() <= costs.max()
    assert sharding.partition_units([5.0], 4) == [[0], [], [], []]


def test_row_slabs_cover_with_halo():
    slabs = sharding.row_slabs(200_000, 8, 16)
    assert slabs[0][0] == 0 and slabs[-1][1] == 200_000
    for (a0, a1, i0, i1), (b0, b1, j0, j1) in zip(slabs[:-1], slabs[1:]):
        assert a1 == b0 and i1 == a1 + 16 and j0 == b0 - 16
    assert slabs[0][2] == 0 and slabs[-1][3] == 200_000
    assert sharding.row_slabs(10, 1, 16) == [(0, 10, 0, 10)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        n = [7, 3][rank]                              # ragged counts; rank 1's buffer is short
        rec = np.zeros(n, dtype=_lib.CANDIDATE_DTYPE)
        rec["row"] = rng.integers(0, 1000, n)
        rec["col"] = rec["row"] + rng.integers(0, 200, n)
        rec["score"] = rng.random(n).astype(np.float32)
        rec["log10p"] = -rng.random(n).astype(np.float32)
        cap = [16, 4][rank]
        buf = torch.zeros((cap, 4), dtype=torch.int32)
        buf[:n] = torch.from_numpy(rec.view(np.int32).reshape(n, 4))
        gathered, counts = sharding.gather_candidates(buf, n)
        merged = sharding.merge_candidates(gathered, counts, row_offsets=[0, 5000])
        q.put((rank, rec, merged))
        # empty rank: nothing to send must still work
        gathered, counts = sharding.gather_candidates(buf, 0 if rank == 0 else n)
        assert counts.tolist() == [0, 3]
    except Exception as e:  # surface the failure instead of a queue timeout
        q.put((rank, None, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_gather_candidates_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] is not None for r in res), [r[2] for r in res]
    local = [res[0][1], res[1][1]]
    expect = local[1].copy()
    expect["row"] += 5000
    expect["col"] += 5000
    for rank in range(2):
        merged = res[rank][2]
        assert len(merged) == 10
        assert np.array_equal(merged[:7], local[0])
        assert np.array_equal(merged[7:], expect)
