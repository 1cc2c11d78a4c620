"""CPU-side tests: host bookkeeping that mirrors the reference, and the C ABI
surface (the library must load and export every symbol of include/*.h)."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN, REPO, load_coo


def test_library_exports_every_declared_symbol():
    from chromosight_b200 import _lib
    header = open(os.path.join(REPO, "include", "chromosight_b200.h")).read()
    declared = set(re.findall(r"\b(cs_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cs_version() == _lib.ABI_VERSION == int(re.search(r"#define CS_ABI_VERSION (\d+)", header).group(1))
    assert lib.cs_launch_count() == 0


def test_layout_arithmetic():
    from chromosight_b200 import _lib
    lib = _lib.load()
    L = _lib.Layout()
    assert lib.cs_layout_band(ctypes.byref(L), 1000, 1000, -32, 233) == 0
    assert L.pitch % 4 == 0 and L.pitch >= 233 + 32
    assert L.n_elems >= 999 * L.pitch + 1000 + 32
    assert lib.cs_layout_dense(ctypes.byref(L), 10, 13) == 0
    assert L.pitch == 16 and L.dense == 1
    assert lib.cs_layout_band(ctypes.byref(L), 10, 10, 5, 2) != 0
    assert b"band" in lib.cs_last_error()


def test_band_gap_plan():
    """cs_pearson_plan reports the gap of zeros a banded image needs for the Pearson kernel to
    skip its alias fix-up; cs_layout_band_padded provides it (host arithmetic only)."""
    from chromosight_b200 import _lib
    lib = _lib.load()
    n, D, k = 200_000, 200, 17
    H = W = n + 2 * (k - 1)
    od_lo, od_hi = 0, D + 2 * k - 1  # image diagonals of the scores of the bench call
    id_lo, id_hi = od_lo - (k - 1), od_hi + (k - 1)
    L = _lib.Layout()
    assert lib.cs_layout_band(ctypes.byref(L), H, W, id_lo, id_hi) == 0
    kern = np.ones((k, k))
    kd = _lib.KernelDesc()
    kd.kh = kd.kw = k
    arr = np.ascontiguousarray(kern.ravel())
    kd.k_corr = kd.k_mask = kd.k2_mask = arr.ctypes.data
    po = _lib.PearsonOpts()
    tr, gap = ctypes.c_int32(0), ctypes.c_int32(-1)
    rc = lib.cs_pearson_plan(ctypes.byref(L), ctypes.byref(kd), ctypes.byref(po), k - 1, k - 1 + n, k - 1, k - 1 + n,
                             od_lo, od_hi, ctypes.byref(tr), ctypes.byref(gap))
    assert rc == 0, lib.cs_last_error()
    assert tr.value == 32 and 0 < gap.value <= 64
    pitch0 = L.pitch
    assert lib.cs_layout_band_padded(ctypes.byref(L), H, W, id_lo, id_hi, gap.value) == 0
    assert L.pitch % 4 == 0 and L.pitch - (id_hi - id_lo) >= gap.value and L.pitch <= pitch0 + gap.value + 4
    assert L.n_elems >= H * (L.pitch + 1)
    # a dense image needs none
    assert lib.cs_layout_dense(ctypes.byref(L), 500, 500) == 0
    rc = lib.cs_pearson_plan(ctypes.byref(L), ctypes.byref(kd), ctypes.byref(po), 8, 492, 8, 492, -483, 483,
                             ctypes.byref(tr), ctypes.byref(gap))
    assert rc == 0 and gap.value == 0


def test_no_silent_cpu_fallback():
    """Without a GPU the hot path must raise, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from chromosight_b200 import _lib
    from chromosight_b200.utils import detection as cud
    sig = sp.random(40, 40, density=0.3, random_state=0, format="csr")
    with pytest.raises(_lib.BackendError):
        cud.normxcorr2(sig, np.arange(9.0).reshape(3, 3))


def test_masks_match_reference():
    from chromosight_b200.utils import preprocessing as cup
    z = np.load(os.path.join(GOLDEN, "preproc_cases.npz"))
    vr = z["mask_valid_rows"]
    m1 = cup.make_missing_mask((12, 12), vr, vr, max_dist=3, sym_upper=True)
    assert m1.dtype == bool and np.array_equal(m1.toarray(), z["mask_m1"])
    assert np.array_equal(cup.frame_missing_mask(m1, (5, 5), True, 3).toarray(), z["mask_f1"])
    m2 = cup.make_missing_mask((12, 9), vr, np.array([0, 2, 3, 4, 6, 8]), sym_upper=False)
    assert np.array_equal(m2.toarray(), z["mask_m2"])
    assert np.array_equal(cup.frame_missing_mask(m2, (5, 3), False, None).toarray(), z["mask_f2"])
    m3 = cup.make_missing_mask((12, 12), vr, vr, max_dist=None, sym_upper=True)
    assert np.array_equal(m3.toarray(), z["mask_m3"])
    assert np.array_equal(cup.frame_missing_mask(m3, (3, 5), True, None).toarray(), z["mask_f3"])
    assert np.array_equal(cup.frame_missing_mask(m1, (3, 7), True, 2).toarray(), z["mask_f4"])
    with pytest.raises(ValueError):
        cup.make_missing_mask((12, 9), vr, vr, sym_upper=True)
    with pytest.raises(ValueError):
        cup.frame_missing_mask(m1.astype(float), (3, 3))
    # reference's worked example (pre:578-585)
    valid = np.array([0, 2, 4])
    exp = np.array([[0, 1, 0, 0, 0], [0, 1, 1, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 1, 1], [0, 0, 0, 0, 0]], bool)
    assert np.array_equal(cup.make_missing_mask((5, 5), valid, valid, max_dist=1, sym_upper=True).toarray(), exp)


def test_small_preprocessing_helpers():
    from chromosight_b200.utils import preprocessing as cup
    z = np.load(os.path.join(GOLDEN, "preproc_cases.npz"))
    raw = load_coo(z, "d1_upper_raw").tocsr()
    assert np.array_equal(cup.diag_trim(raw, 17).toarray(), load_coo(z, "trim_out").toarray())
    assert np.allclose(cup.ztransform(raw.tocoo()).toarray(), load_coo(z, "zt_out").toarray())
    from chromosight_b200 import kernels
    u, v = cup.factorise_kernel(kernels.loops["kernels"][0], 0.999)
    assert np.allclose(u @ v, z["fact_loops_uv"], atol=1e-12)
    # zero_pad_sparse docstring example (pre:655-661)
    m = sp.csr_matrix(np.array([[1, 2], [10, 20]]))
    exp = np.array([[0, 0, 0, 0, 0, 0], [0, 0, 1, 2, 0, 0], [0, 0, 10, 20, 0, 0], [0, 0, 0, 0, 0, 0]])
    assert np.array_equal(cup.zero_pad_sparse(m, 2, 1).toarray(), exp)
    assert np.array_equal(cup.valid_to_missing(np.array([0, 2]), 4), [1, 3])
    with pytest.raises(ValueError):
        cup.check_missing_mask(sp.csr_matrix(np.eye(3)), sp.csr_matrix(np.eye(3, dtype=bool)))
    cup.check_missing_mask(sp.csr_matrix(np.eye(3)), sp.csr_matrix(np.eye(3, k=1, dtype=bool)))


def test_stats_known_answers():
    """BH q-values vs R's p.adjust (reference tests/test_stats.py)."""
    from chromosight_b200.utils import stats as cus
    from oracle import pearson_oracle as po
    assert np.allclose(cus.fdr_correction(np.array([0.1, 0.1, 0.05, 0.01])), [0.1, 0.1, 0.1, 0.04])
    r = np.array([0.0, 0.3, -0.7, 1.0])
    assert np.allclose(cus.corr_to_pval(r, 289), po.corr_to_log10_pval(r, 289))
    assert cus.corr_to_pval(np.array([1.0]), 49)[0] == -np.inf


def test_foci_match_reference():
    from chromosight_b200.utils import detection as cud
    z = np.load(os.path.join(GOLDEN, "foci_cases.npz"))
    corr = load_coo(z, "corr")
    for thr in (20, 30):
        coords, lab = cud.pick_foci(corr.copy(), thr / 100)
        assert np.array_equal(coords, z[f"thr{thr}_coords"])
        assert np.array_equal(lab.toarray(), load_coo(z, f"thr{thr}_labels").toarray())
    assert cud.pick_foci(corr, 2.0) == (None, None)


def test_label_foci_known_answer():
    """Exact labels of the reference's test_label_spec (tests/test_detection.py:204-238)."""
    from chromosight_b200.utils import detection as cud
    m = np.array([
        [1, 0, 1, 1, 0],
        [1, 0, 0, 0, 1],
        [1, 0, 1, 0, 1],
        [0, 0, 1, 1, 0],
        [1, 0, 0, 0, 1],
    ])
    num, lab = cud.label_foci(sp.coo_matrix(m))
    assert num == 6
    exp = np.array([
        [1, 0, 2, 2, 0],
        [1, 0, 0, 0, 3],
        [1, 0, 4, 0, 3],
        [0, 0, 4, 4, 0],
        [5, 0, 0, 0, 6],
    ])
    assert np.array_equal(lab.toarray(), exp)
    n2, filt = cud.filter_foci(lab.copy(), min_size=3)
    assert n2 == 2 and set(np.unique(filt.data)) == {1, 4}


def test_resize_and_crop_kernel(presets):
    """Sizes and centre of resize_kernel / crop_kernel as pinned by the reference's
    tests/test_preprocessing.py:118-180 (odd sizes, min_size, centre preserved)."""
    from chromosight_b200.utils import preprocessing as cup
    K = presets.loops["kernels"][0]
    assert cup.resize_kernel(K, factor=9 / 17).shape == (9, 9)
    assert cup.resize_kernel(K, factor=0.1).shape == (7, 7)          # min_size
    assert cup.resize_kernel(K, kernel_res=10000, signal_res=5000).shape[0] % 2 == 1
    up = cup.resize_kernel(K, factor=2.0)
    assert up.shape[0] % 2 == 1 and np.isclose(up[up.shape[0] // 2, up.shape[1] // 2], K[8, 8])
    with pytest.raises(ValueError):
        cup.resize_kernel(K[:, :15], factor=1)
    with pytest.raises(ValueError):
        cup.resize_kernel(K[:16, :16], factor=1)
    with pytest.raises(ValueError):
        cup.resize_kernel(K, factor=1, kernel_res=1)
    assert cup.crop_kernel(K, (7, 7)).shape == (7, 7)
    assert cup.crop_kernel(K, (8, 8)).shape == (9, 9)
    assert cup.crop_kernel(K, (21, 5)).shape == (17, 5)
    assert np.array_equal(cup.crop_kernel(K, (7, 7)), K[5:12, 5:12])


def test_global_host_filters_match_reference():
    """remove_neighbours (det:348-384), pileup_patterns (det:158-174) and fdr_correction
    (stats:7-40) against outputs of the unmodified reference (make_golden_hostfilters.py)."""
    import pandas as pd
    from chromosight_b200.utils import detection as cud, stats as cus
    z = np.load(os.path.join(GOLDEN, "hostfilters.npz"))
    i = 0
    while f"rn{i}_mask" in z.files:
        t = pd.DataFrame({"bin1": z[f"rn{i}_bin1"], "bin2": z[f"rn{i}_bin2"], "score": z[f"rn{i}_score"]})
        got = cud.remove_neighbours(t, win_size=int(z[f"rn{i}_win"]))
        exp = z[f"rn{i}_mask"]
        # equal scores are ordered arbitrarily by the reference's quicksort: the kept SET may differ
        # there, the number kept and the separation property may not
        if len(np.unique(t.score)) == len(t):
            assert np.array_equal(got, exp)
        else:
            assert abs(int(got.sum()) - int(exp.sum())) <= max(2, len(t) // 50)
        kept = t[got]
        d1 = np.abs(kept.bin1.values[:, None] - kept.bin1.values[None, :])
        d2 = np.abs(kept.bin2.values[:, None] - kept.bin2.values[None, :])
        w = int(z[f"rn{i}_win"])
        assert ((d1 < w) & (d2 < w)).sum() == len(kept)
        i += 1
    assert i == 4
    assert np.allclose(cud.pileup_patterns(z["pile_in"]), z["pile_out"], equal_nan=True)
    assert np.allclose(cus.fdr_correction(z["fdr_in"]), z["fdr_out"])


@pytest.mark.parametrize("threads", [1, 4])
def test_expand_rows_narrow_format(threads):
    """Host side of the narrow result format (csrc/host_expand.cpp): float32 score / log10 p and
    uint8 diagonal offsets -> float64 data and int32 column indices, bit for bit."""
    from chromosight_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(threads)
    rows, dlo = 5000, -3
    counts = rng.integers(0, 240, size=rows)
    counts[rng.random(rows) < 0.1] = 0
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    n = int(indptr[-1])
    score = rng.standard_normal(n).astype(np.float32)
    logp = (-np.abs(rng.standard_normal(n)) * 5).astype(np.float32)
    logp[::97] = -np.inf
    off = np.concatenate([np.sort(rng.choice(256, size=c, replace=False)) for c in counts]).astype(np.uint8)
    for r0, r1 in ((0, rows), (123, 4001)):
        data = np.full(n + 3, 7.0)[3:]           # deliberately not 64-byte aligned
        lp = np.full(n, 7.0)
        idx = np.full(n, -1, dtype=np.int32)
        idx2 = np.full(n, -1, dtype=np.int32)
        rc = lib.cs_expand_rows(score.ctypes.data, logp.ctypes.data, off.ctypes.data, indptr.ctypes.data,
                                r0, r1, dlo, data.ctypes.data, lp.ctypes.data, idx.ctypes.data,
                                idx2.ctypes.data, threads)
        assert rc == 0
        a, b = int(indptr[r0]), int(indptr[r1])
        row_of = np.repeat(np.arange(rows), counts)
        assert np.array_equal(data[a:b], score[a:b].astype(np.float64))
        assert np.array_equal(lp[a:b], logp[a:b].astype(np.float64))
        exp_idx = (row_of + dlo + off.astype(np.int64)).astype(np.int32)
        assert np.array_equal(idx[a:b], exp_idx[a:b]) and np.array_equal(idx2[a:b], exp_idx[a:b])
        # nothing outside the row range is touched
        assert np.all(data[:a] == 7.0) and np.all(data[b:] == 7.0) and np.all(idx[:a] == -1) and np.all(idx[b:] == -1)
    # without p-values / second index array
    data = np.zeros(n)
    idx = np.zeros(n, dtype=np.int32)
    assert lib.cs_expand_rows(score.ctypes.data, None, off.ctypes.data, indptr.ctypes.data, 0, rows, dlo,
                              data.ctypes.data, None, idx.ctypes.data, None, threads) == 0
    assert np.array_equal(data, score.astype(np.float64))


def test_cli_fixture_reproduces_reference_held_answers():
    """tests/golden/cli_example.npz (the `chromosight detect` chain driven through the unmodified
    reference's functions, make_golden_cli.py) against the answers the reference's repository
    holds: 89 patterns of `chromosight test` (cli:185-199) and the 59 / 57 / 55 rows of
    docs/notebooks/detect/example_*.tsv -- same coordinates, scores to 1e-9."""
    z = np.load(os.path.join(GOLDEN, "cli_example.npz"), allow_pickle=False)
    assert len(z["loops_default_bin1"]) == 89
    for case, n in (("loops_nb", 59), ("borders", 57), ("hairpins", 55)):
        assert len(z[f"{case}_bin1"]) == n == len(z[f"{case}_held_bin1"])
        o = np.lexsort((z[f"{case}_bin2"], z[f"{case}_bin1"]))
        h = np.lexsort((z[f"{case}_held_bin2"], z[f"{case}_held_bin1"]))
        assert np.array_equal(z[f"{case}_bin1"][o], z[f"{case}_held_bin1"][h])
        assert np.array_equal(z[f"{case}_bin2"][o], z[f"{case}_held_bin2"][h])
        assert np.abs(z[f"{case}_score"][o] - z[f"{case}_held_score"][h]).max() < 1e-9


def test_distance_law_custom_fun_host_reduction():
    """distance_law(fun=...) with a reduction other than nanmean (pre:129-136): grouped on the
    host; checked against a dense restatement of pre:178-188 (and the live reference when
    oracle/_ref is present)."""
    from chromosight_b200 import synthetic
    from chromosight_b200.utils import preprocessing as cup
    from oracle import ref_loader
    raw, detect = synthetic.band_counts(300, 45, seed=6, missing_frac=0.05, max_dist=40)
    A = raw.toarray()
    ok = np.zeros(300, bool)
    ok[detect] = True
    ref = ref_loader.load()
    for fun in (np.nanmedian, np.nanmax):
        law, n_diags = cup._law_host_fun(raw.tocsr(), detect, 40, fun)
        assert n_diags == 41 and np.all(law[41:] == 0)
        for d in range(41):
            v = np.diagonal(A, d)[ok[: 300 - d] & ok[d:]]
            v = v[v > 0]
            assert (np.isnan(law[d]) and len(v) == 0) or np.isclose(law[d], fun(v))
        if ref is not None:
            exp = ref[1].distance_law(raw.tocsr(), detectable_bins=detect, max_dist=40, smooth=False, fun=fun)
            assert np.allclose(law[:41], exp[:41], equal_nan=True)
