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


# --------------------------------------------------------------------------- drivers, world size 2
class _FakeMap:
    """ContactMap stand-in: no file, no GPU."""

    def __init__(self, extent, inter):
        self.extent, self.inter, self.matrix = extent, inter, None

    def create_mat(self):
        self.matrix = "loaded"

    def destroy_mat(self):
        self.matrix = None


class _FakeClr:
    binsize = 1000
    chromnames = ["a", "b", "c"]
    _off = {"a": (0, 400), "b": (400, 700), "c": (700, 800)}
    shape = (800, 800)

    def extent(self, c):
        return self._off[c]


class _FakeGenome:
    def __init__(self):
        import pandas as pd
        self.clr = _FakeClr()
        self.max_dist = 50
        rows = []
        for i, c1 in enumerate(self.clr.chromnames):
            for j, c2 in enumerate(self.clr.chromnames):
                if j >= i:
                    rows.append({"chr1": c1, "chr2": c2,
                                 "contact_map": _FakeMap([self.clr.extent(c1), self.clr.extent(c2)], i != j)})
        self.sub_mats = pd.DataFrame(rows)
        start = np.arange(800) * 1000
        self.bins = pd.DataFrame({"chrom": np.repeat(["a", "b", "c"], [400, 300, 100]),
                                  "start": start, "end": start + 1000})

    def get_full_mat_pattern(self, c1, c2, t):
        t = t.copy()
        t["bin1"] += self.clr.extent(c1)[0]
        t["bin2"] += self.clr.extent(c2)[0]
        return t

    def bins_to_coords(self, idx):
        return self.bins.iloc[np.asarray(idx), :][["chrom", "start", "end"]]


def _fake_detector(cm, cfg, kernel, coords=None, full=False, tsvd=None, dump=None):
    """Deterministic per-sub-matrix table: stands in for the GPU pattern_detector."""
    import pandas as pd
    assert cm.matrix == "loaded"
    (s1, e1), (s2, e2) = cm.extent
    n = (e1 - s1 + e2 - s2) % 7
    if n == 0:
        return None, None
    b1 = (np.arange(n) * 37) % (e1 - s1)
    b2 = np.minimum(b1 + 20 + np.arange(n), e2 - s2 - 1) if not cm.inter else (np.arange(n) * 11) % (e2 - s2)
    t = pd.DataFrame({"bin1": b1, "bin2": b2, "score": 0.3 + 0.01 * np.arange(n) + 0.001 * s1,
                      "pvalue": 1e-3 / (1 + np.arange(n))})
    return t, np.full((n, 3, 3), float(s1 + s2))


def _driver_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from chromosight_b200 import driver
        from chromosight_b200.utils import detection as cud
        cud.pattern_detector = _fake_detector
        hg = _FakeGenome()
        cfg = {"kernels": [np.ones((3, 3)), np.eye(3)], "max_iterations": 2, "min_separation": 5000,
               "min_dist": 0, "max_dist": 50000}
        table, wins = driver.detect(hg, cfg, full=True)
        q.put((rank, table.to_dict("list"), wins.tolist()))
    except Exception as e:
        q.put((rank, None, repr(e)))
        raise
    finally:
        if world > 1:
            dist.destroy_process_group()


def _run_driver(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_driver_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(r[1] is not None for r in res), [r[2] for r in res]
    return res


def test_detect_driver_world2_gloo_matches_single_process():
    """The sharded detect loop (driver.detect: LPT partition, all_gather_object of the tables,
    global filters) gives every rank the table a single process computes."""
    single = _run_driver(1)[0]
    both = _run_driver(2)
    assert len(single[1]["bin1"]) > 0
    for r in both:
        assert r[1] == single[1]
        assert r[2] == single[2]


# --------------------------------------------------------------------------- row slabs (strong scaling)
def _slab_worker(rank, world, port, q):
    """One rank of the row-slab path with the oracle standing in for the device kernels: partial
    distance-law sums of the owned rows -> ONE all-reduce -> slab detrend -> Pearson map of the
    slab's square sub-matrix -> candidates of the owned rows -> fixed-size gather."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scipy.sparse as sp
        from chromosight_b200 import kernels, rowslab, synthetic
        from chromosight_b200.utils import preprocessing as cup
        from oracle import pearson_oracle as po
        kernel = kernels.loops_small["kernels"][0]
        k, n, D, thr = kernel.shape[0], 260, 24, 0.3
        raw, detect = synthetic.band_counts(n, D + k, seed=4, missing_frac=0.04, max_dist=D)
        raw = raw.tocsr()
        plan = rowslab.slab_plan(n, world, k, D)[rank]
        r0, r1, in0, in1 = plan
        # partial law sums of the owned rows (what K0a accumulates on the device)
        own = rowslab.owned_rows(raw, r0, r1).toarray()
        ok = np.zeros(n, bool)
        ok[detect] = True
        nd = min(n, D + k + 1)
        psum, pcnt = np.zeros(nd), np.zeros(nd)
        for d in range(nd):
            v = np.diagonal(own, d)[ok[: n - d] & ok[d:]]
            v = v[v > 0]
            psum[d], pcnt[d] = v.sum(), len(v)
        buf = torch.from_numpy(np.concatenate([psum, pcnt]))
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        law = rowslab.law_from_sums(buf[:nd].numpy(), buf[nd:].numpy(), n)
        ref_law = po.distance_law_dense(raw.toarray(), detect, D + k)
        ref_law[np.isnan(ref_law)] = 0
        assert np.allclose(law, ref_law, rtol=1e-13)
        # slab inputs (host arithmetic of slab_inputs, division by the global law)
        sub = raw[in0:in1, in0:in1].toarray()
        rr, cc = np.indices(sub.shape)
        with np.errstate(all="ignore"):
            mat = np.where(sub != 0, sub / law[np.abs(rr - cc)], 0.0)
        mat[mat >= 10] = 1.0
        mat = np.triu(np.tril(mat, D + k))
        mat[np.isnan(mat)] = 0
        det = np.asarray(detect)
        det = det[(det >= in0) & (det < in1)] - in0
        mask = po.make_missing_mask_dense(mat.shape, det, det, max_dist=D, sym_upper=True)
        kw = dict(max_dist=D, sym_upper=True, full=True, missing_tol=0.5)
        r, _ = po.normxcorr2_dense(mat, kernel, missing_mask=mask, **kw)
        r = np.triu(np.tril(r, D))
        ys, xs = np.nonzero((r >= thr) & (r != 0))
        rec = np.zeros(len(ys), dtype=_lib.CANDIDATE_DTYPE)
        rec["row"], rec["col"], rec["score"] = ys, xs, r[ys, xs]
        mine = rowslab.owned_candidates(rec, plan)
        cap = 4096
        sendbuf = torch.zeros((cap, 4), dtype=torch.int32)
        sendbuf[: len(mine)] = torch.from_numpy(mine.view(np.int32).reshape(-1, 4))
        gathered, counts = sharding.gather_candidates(sendbuf, len(mine), cap=cap)
        merged = rowslab.merge_sorted([sharding.merge_candidates(gathered, counts)])
        q.put((rank, merged))
    except Exception as e:
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_row_slab_path_world2_gloo():
    """Row slabs of ONE chromosome over two ranks give the candidate pixels of a single run
    over the whole map (the law all-reduce, the halos and the merge are what is tested; the
    device kernels are replaced by the oracle)."""
    import scipy.sparse as sp
    from chromosight_b200 import kernels, rowslab, synthetic
    from oracle import pearson_oracle as po
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert not isinstance(res[0][1], str), res[0][1]
    # single run over the whole map
    kernel = kernels.loops_small["kernels"][0]
    k, n, D, thr = kernel.shape[0], 260, 24, 0.3
    raw, detect = synthetic.band_counts(n, D + k, seed=4, missing_frac=0.04, max_dist=D)
    mat = np.triu(np.tril(po.detrend_dense(raw.toarray(), detect, D + k, 10), D + k))
    mat[np.isnan(mat)] = 0
    mask = po.make_missing_mask_dense(mat.shape, detect, detect, max_dist=D, sym_upper=True)
    r, _ = po.normxcorr2_dense(mat, kernel, missing_mask=mask, max_dist=D, sym_upper=True, full=True,
                               missing_tol=0.5)
    r = np.triu(np.tril(r, D))
    ys, xs = np.nonzero((r >= thr) & (r != 0))
    for rank in range(2):
        merged = res[rank][1]
        assert np.array_equal(merged["row"], ys) and np.array_equal(merged["col"], xs)
        assert np.allclose(merged["score"], r[ys, xs].astype(np.float32), atol=1e-6)
    assert len(ys) > 20
    # foci of the merged pixels = foci of the whole map
    coords = rowslab.foci_of_candidates(res[0][1], r.shape, thr)
    from chromosight_b200.utils.detection import pick_foci
    coords0, _ = pick_foci(sp.coo_matrix(r), thr)
    assert np.array_equal(coords, coords0)
    # plan arithmetic
    plan = rowslab.slab_plan(200_000, 8, 17, 200)
    assert plan[0][:3] == (0, 25_000, 0) and plan[-1][1] == plan[-1][3] == 200_000
    assert all(p[2] == p[0] - 17 for p in plan[1:]) and all(p[3] == p[1] + 200 + 51 for p in plan[:-1])
