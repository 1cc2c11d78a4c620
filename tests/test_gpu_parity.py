"""CUDA path vs the reference's golden outputs and vs the oracle (run on a B200:
`pytest -m gpu`).  Everything goes through the C ABI (cs_normxcorr2_host & co).

Tolerances (BASELINE.json north_star): Pearson scores within 1e-5 absolute;
log10 p-values within 1e-4 relative where the score itself is significant.
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN, golden_case_names, load_case, load_coo

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-5


def _score_tol(cond):
    """1e-5 wherever a window is reasonably conditioned (rms/std of its present
    pixels <= 30: every window of a noisy contact map).  The signal is stored in
    float32 on the device, so on synthetic, nearly constant windows the
    reachable accuracy degrades like eps32 * rms/std; exactly flat windows
    (cond = inf, score = round-off noise in the reference) are held to 1e-5."""
    if cond is None:
        return SCORE_TOL
    c = np.where(np.isfinite(cond), cond, 1.0)
    return SCORE_TOL * np.maximum(1.0, c / 30.0)


def _compare(r, p, corr_ref, pval_ref, n_obs=None, cond=None, strict=False):
    """Scores within 1e-5 of the reference everywhere; p-values (a) equal to the
    reference's formula evaluated at the returned score (2e-6: float32 evaluation), which
    pins the device p-value code, and (b) within 1e-4 relative of the reference's own
    values over the score range where foci live (log10 p is ill-conditioned in r
    as |r| -> 1, so (b) is restricted to 0.1 <= |r| <= 0.9)."""
    from oracle import pearson_oracle as po
    r = r.toarray() if sp.issparse(r) else np.asarray(r)
    assert r.shape == corr_ref.shape
    d = np.abs(r - corr_ref)
    tol = SCORE_TOL if strict else _score_tol(cond)
    bad = d > tol
    assert not bad.any(), f"{int(bad.sum())} windows off, max |dr| = {d.max():.3e} at {np.unravel_index(np.argmax(d), d.shape)}"
    if pval_ref is not None:
        assert p is not None
        p = p.toarray() if sp.issparse(p) else np.asarray(p)
        nz = r != 0
        if n_obs is not None:
            exp = po.corr_to_log10_pval(r[nz], n_obs[nz])
            got = p[nz]
            fin = np.isfinite(exp)
            assert np.array_equal(np.isneginf(got), np.isneginf(exp))
            assert np.array_equal(np.isnan(got), np.isnan(exp))
            # the device evaluates the formula in float32 (log10 erfcx(a) - a^2 log10 e)
            assert np.allclose(got[fin], exp[fin], rtol=2e-6, atol=2e-6)
        sel = nz & (np.abs(corr_ref) >= 0.1) & (np.abs(corr_ref) <= 0.9) & np.isfinite(pval_ref)
        assert np.allclose(p[sel], pval_ref[sel], rtol=1e-4, atol=1e-4)
        assert np.all(p[~nz] == 0)


@pytest.mark.parametrize("name", golden_case_names())
def test_normxcorr2_golden(name):
    from chromosight_b200.utils import detection as cud
    from oracle import pearson_oracle as po
    signal, kernel, kw, dense, corr_ref, pval_ref = load_case(name)
    sig = signal.toarray() if dense else signal
    r, p = cud.normxcorr2(sig, kernel, **kw)
    okw = dict(kw)
    mask = okw.pop("missing_mask", None)
    _, _, n_obs, cond = po.normxcorr2_dense(signal.toarray(), kernel, return_cond=True,
                                            missing_mask=None if mask is None else mask.toarray(), **okw)
    if dense:
        assert isinstance(r, np.ndarray)
    else:
        assert sp.issparse(r) and r.format == "csr" and r.dtype == np.float64
        assert r.nnz == 0 or np.all(r.data != 0)
    # every fixture, including the synthetic smooth ones, meets 1e-5 on every window
    _compare(r, p, corr_ref, pval_ref, n_obs, cond, strict=True)


def test_xcorr2_golden():
    from chromosight_b200.utils import detection as cud
    z = np.load(os.path.join(GOLDEN, "xcorr2_cases.npz"))
    for nm in ("gauss", "band", "band_const", "band_tsvd", "band_rect"):
        sig = load_coo(z, f"{nm}_signal").tocsr()
        kw = json.loads(str(z[f"{nm}_kwargs"]))
        out = cud.xcorr2(sig, z[f"{nm}_kernel"], **kw)
        ref = load_coo(z, f"{nm}_out").toarray()
        scale = max(1.0, np.abs(ref).max())
        # raw sums of up to 289 products: relative tolerance on the map's scale
        assert np.abs(out.toarray() - ref).max() <= 2e-5 * scale, nm


def test_xcorr2_dense_signal():
    """The reference's own test of the dense branch (tests/test_detection.py:241-270,
    _xcorr2_dense det:726-804): a dense (ndarray / np.matrix) signal returns an ndarray equal
    to the sparse result and to scipy's correlate2d thresholded at 1e-4, with its maximum at
    the mode of the 2-D normal the signal holds."""
    import scipy.signal as sig
    from scipy.stats import multivariate_normal
    from chromosight_b200.utils import detection as cud

    def gauss_mat(meanx, meany, std, shape=(100, 100)):
        k = multivariate_normal(mean=(meanx, meany), cov=np.eye(2) * std)
        x, y = np.linspace(-10, 10, shape[0]), np.linspace(-10, 10, shape[1])
        xx, yy = np.meshgrid(x, y)
        return k.pdf(np.c_[xx.ravel(), yy.ravel()]).reshape(shape)

    gauss_kernel = gauss_mat(0, 0, 5.0, shape=(7, 7))
    n_checked = 0
    for mx, my, sd in ((-1.5, -1.0, 0.3), (-1.0, 0.5, 0.9), (0.0, 1.0, 1.8), (-0.5, 1.0, 2.7)):
        dense = gauss_mat(mx, my, sd)
        exp_row, exp_col = np.where(dense == dense.max())
        out_sparse = cud.xcorr2(sp.coo_matrix(dense), gauss_kernel, threshold=1e-4)
        assert sp.issparse(out_sparse)
        for signal in (dense, np.asmatrix(dense)):          # ndarray and what .todense() returns
            out = cud.xcorr2(signal, gauss_kernel, threshold=1e-4)
            assert isinstance(out, np.ndarray) and not sp.issparse(out) and out.shape == dense.shape
            obs_row, obs_col = np.where(out == out.max())
            assert np.all(np.isin(obs_row, exp_row)) and np.all(np.isin(obs_col, exp_col))
            ref = np.zeros(dense.shape)
            ref[3:-3, 3:-3] = sig.correlate2d(dense, gauss_kernel, "valid")
            ref[ref < 1e-4] = 0
            near = np.abs(np.abs(ref) - 1e-4) < 1e-8           # the threshold itself is float32-fuzzy
            assert np.abs(out - ref)[~near].max() <= 2e-5 * max(1.0, np.abs(ref).max())
            assert np.abs(out - out_sparse.toarray()).max() == 0
            n_checked += 1
    assert n_checked == 8


def test_detrend_golden():
    from chromosight_b200.utils import preprocessing as cup
    z = np.load(os.path.join(GOLDEN, "preproc_cases.npz"))
    for i in range(2):
        for tag in ("upper", "sym"):
            raw = load_coo(z, f"d{i}_{tag}_raw").tocsr()
            det = z[f"d{i}_{tag}_detect"]
            D = int(z[f"d{i}_{tag}_max_dist"])
            law = cup.distance_law(raw, detectable_bins=det, max_dist=D, smooth=False)
            assert np.allclose(law, z[f"d{i}_{tag}_law"], rtol=1e-12, equal_nan=True)
            out = cup.detrend(raw, detectable_bins=det, max_dist=D, max_val=10)
            ref = load_coo(z, f"d{i}_{tag}_out").toarray()
            assert sp.issparse(out) and out.format == "csr"
            assert np.allclose(out.toarray(), ref, rtol=1e-12)
            out = cup.detrend(raw, detectable_bins=det, max_dist=D, max_val=None)
            ref = load_coo(z, f"d{i}_{tag}_out_nomax").toarray()
            assert np.allclose(out.toarray(), ref, rtol=1e-12)


@pytest.fixture
def mask_form(request):
    """Run a test with the mask reaching the device in one of its forms: "auto" (a mask made
    by this package: geometry attached), "foreign" (the same pixels without the tag, as the
    reference's make_missing_mask returns them: recognised from the pattern), "pixels"
    (geometry ignored: NaN sentinels + per-pixel exact path)."""
    from chromosight_b200.utils import detection as cud
    form = request.param
    old = cud.MASK_FORM
    cud.MASK_FORM = "pixels" if form == "pixels" else "auto"
    yield form
    cud.MASK_FORM = old


def _as_form(mask, form):
    return mask.copy() if form == "foreign" else mask


@pytest.mark.parametrize("mask_form", ["auto", "foreign", "pixels"], indirect=True)
@pytest.mark.parametrize("seed,n,D,kname,tol", [
    (21, 700, 60, "loops", 0.5), (22, 513, 25, "loops_small", 0.5),
    (23, 400, 90, "hairpins", 0.75), (24, 333, 40, "borders", 0.75),
])
def test_production_call_vs_oracle(seed, n, D, kname, tol, presets, mask_form):
    """pattern_detector's call (det:242-263) on seeded synthetic maps, checked
    against the oracle on every window, strictly to 1e-5."""
    from chromosight_b200 import synthetic
    from chromosight_b200.utils import detection as cud, preprocessing as cup
    from oracle import pearson_oracle as po
    kernel = getattr(presets, kname)["kernels"][0]
    k = kernel.shape[0]
    raw, detect = synthetic.band_counts(n, D + k, seed=seed, missing_frac=0.04, max_dist=D)
    mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
    mat = cup.diag_trim(mat.tocsr(), D + k)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    mask = _as_form(cup.make_missing_mask(mat.shape, detect, detect, max_dist=D, sym_upper=True), mask_form)
    assert hasattr(mask, "_cs_geometry") == (mask_form != "foreign")
    kw = dict(max_dist=D, sym_upper=True, full=True, missing_tol=tol, pval=True)
    r, p = cud.normxcorr2(mat, kernel, missing_mask=mask, **kw)
    r0, p0, nob = po.normxcorr2_dense(mat.toarray(), kernel, missing_mask=mask.toarray(),
                                      return_nobs=True, **kw)
    _compare(r, p, r0, p0, nob, strict=True)
    # the extension that skips scores beyond max_dist equals a diag_trim of the full
    # result (same windows; another tiling, hence another float32 rounding)
    rt, _ = cud.normxcorr2(mat, kernel, missing_mask=mask, trim_to_max_dist=True, **kw)
    full_trim = cup.diag_trim(r.tocsr(), D)
    assert np.array_equal(rt.toarray() != 0, full_trim.toarray() != 0)
    assert np.abs(rt - full_trim).max() <= 5e-6


@pytest.mark.parametrize("mask_form", ["auto", "foreign", "pixels"], indirect=True)
def test_inter_and_odd_shapes_vs_oracle(presets, mask_form):
    from chromosight_b200 import synthetic
    from chromosight_b200.utils import detection as cud, preprocessing as cup
    from oracle import pearson_oracle as po
    kernel = presets.loops["kernels"][0][3:14, :]           # 11 x 17 rectangle
    imat, (vr, vc) = synthetic.inter_counts(301, 187, seed=31, density=0.2, missing_frac=0.05)
    mask = _as_form(cup.make_missing_mask(imat.shape, vr, vc, sym_upper=False), mask_form)
    for full in (True, False):
        kw = dict(max_dist=None, sym_upper=False, full=full, missing_tol=0.6, pval=True)
        r, p = cud.normxcorr2(imat, kernel, missing_mask=mask, **kw)
        r0, p0, nob = po.normxcorr2_dense(imat.toarray(), kernel, missing_mask=mask.toarray(),
                                          return_nobs=True, **kw)
        _compare(r, p, r0, p0, nob, strict=True)
    r, p = cud.normxcorr2(imat, kernel, full=True, pval=True)
    r0, p0 = po.normxcorr2_dense(imat.toarray(), kernel, full=True, pval=True)
    _compare(r, p, r0, p0, strict=True)


@pytest.mark.parametrize("kshape", [(11, 17), (17, 11), (7, 7)])
def test_sym_upper_non_square_kernels_vs_oracle(presets, kshape):
    """sym_upper with non-square kernels: the reference takes sp.triu of the framed map before
    cropping it (det:1098-1099, 1124-1129), so the kept diagonals shift by nk - mk."""
    from chromosight_b200 import synthetic
    from chromosight_b200.utils import detection as cud, preprocessing as cup
    from oracle import pearson_oracle as po
    full_k = presets.loops["kernels"][0]
    r0_, c0_ = (17 - kshape[0]) // 2, (17 - kshape[1]) // 2
    kernel = full_k[r0_:r0_ + kshape[0], c0_:c0_ + kshape[1]]
    n, D = 300, 40
    raw, detect = synthetic.band_counts(n, D + 17, seed=41, missing_frac=0.04, max_dist=D)
    mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + 17, max_val=10)
    mat = cup.diag_trim(mat.tocsr(), D + 17)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    mask = cup.make_missing_mask(mat.shape, detect, detect, max_dist=D, sym_upper=True)
    for use_mask in (True, False):
        kw = dict(max_dist=D, sym_upper=True, full=True, missing_tol=0.6, pval=True)
        mm = mask if use_mask else None
        if use_mask and kshape[0] > kshape[1]:
            # a kernel taller than wide: the max(mk, nk) sub-diagonals of the FRAMED map that
            # frame_missing_mask flags (pre:483-497) reach the signal's main diagonals, and
            # check_missing_mask refuses the call (pre:501-532) -- in the reference as here
            with pytest.raises(ValueError):
                cud.normxcorr2(mat, kernel, missing_mask=mm, **kw)
            continue
        r, p = cud.normxcorr2(mat, kernel, missing_mask=mm, **kw)
        r0, p0, nob = po.normxcorr2_dense(mat.toarray(), kernel, return_nobs=True,
                                          missing_mask=None if mm is None else mm.toarray(), **kw)
        _compare(r, p, r0, p0, nob, strict=True)


def test_error_behaviour(presets):
    """ValueErrors of det:871-889 and pre:520-532."""
    from chromosight_b200.utils import detection as cud
    k = presets.loops_small["kernels"][0]
    sig = sp.random(50, 50, density=0.3, random_state=1, format="csr")
    with pytest.raises(ValueError):
        cud.normxcorr2(sig, sp.csr_matrix(k))
    with pytest.raises(ValueError):
        cud.normxcorr2(sig, np.ones((7, 7)))
    with pytest.raises(ValueError):
        cud.normxcorr2(sig, k, missing_mask=np.zeros((50, 50), bool))
    with pytest.raises(ValueError):
        cud.normxcorr2(sig, k, missing_mask=sp.csr_matrix((50, 50), dtype=float))
    with pytest.raises(ValueError):
        cud.normxcorr2(sig, k, missing_mask=sp.csr_matrix((40, 50), dtype=bool))
    with pytest.raises(ValueError):
        cud.normxcorr2(sp.csr_matrix(np.ones((5, 5))), k, missing_mask=sp.csr_matrix((5, 5), dtype=bool))
    bad = sp.csr_matrix(np.ones((50, 50), dtype=bool))      # everything missing, signal non-zero
    with pytest.raises(ValueError):
        cud.normxcorr2(sig, k, missing_mask=bad, full=True)
    # empty signal -> empty result, not an error
    r, p = cud.normxcorr2(sp.csr_matrix((50, 50)), k, pval=True)
    assert r.nnz == 0 and p.nnz == 0


def test_scale_and_shift_properties_large(presets):
    """Size-independent properties on a map far larger than the oracle can take:
    Pearson scores are invariant to a positive rescaling of the signal, and a map
    made of two identical halves yields identical scores in both halves."""
    from chromosight_b200 import synthetic
    from chromosight_b200.utils import detection as cud
    kernel = presets.loops["kernels"][0]
    n, D = 20000, 120
    raw, _ = synthetic.band_counts(n, D, seed=5, missing_frac=0.0)
    half = raw[: n // 2, : n // 2]
    two = sp.block_diag([half, half], format="csr")
    r, _ = cud.normxcorr2(two, kernel, max_dist=D, sym_upper=True, full=True)
    r2, _ = cud.normxcorr2(two * 8.0, kernel, max_dist=D, sym_upper=True, full=True)
    assert r.nnz > 1e6
    # x8 is exact in floating point except through the 1e-4 thresholds of xcorr2
    assert np.abs((r - r2)).max() <= 1e-6
    h = n // 2
    a = r[: h, : h].toarray() if False else r[200: h - 200, :][:, 200: h - 200]
    b = r[h + 200: n - 200, :][:, h + 200: n - 200]
    assert np.abs((a - b)).max() <= 2e-6


@pytest.mark.parametrize("kname,ksize,tol,pearson", [("loops", 17, 0.5, 0.3), ("borders", 9, 0.75, 0.15)])
def test_full_size_map_against_oracle_crops(kname, ksize, tol, pearson, presets):
    """BASELINE.json's full-size intra map (200k bins, max_dist 200): the oracle cannot
    take the whole map, but a window only sees its own pixels, so random crops of the
    full-size result must equal the oracle run on the cropped signal and mask
    (interior windows); plus candidate thresholding against the host version.
    loops 17x17 = the metric configuration; borders resized to 9x9 = config 3."""
    from chromosight_b200 import synthetic
    from chromosight_b200.session import Session, records_to_numpy
    from chromosight_b200.utils import preprocessing as cup
    from oracle import pearson_oracle as po
    kernel = getattr(presets, kname)["kernels"][0]
    if ksize != kernel.shape[0]:
        kernel = cup.resize_kernel(kernel, factor=ksize / kernel.shape[0])
    k = kernel.shape[0]
    assert k == ksize
    n, D = 200_000, 200
    raw, detect = synthetic.band_counts(n, D + k, seed=11, missing_frac=0.02, max_dist=D)
    mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
    mat = cup.diag_trim(mat.tocsr(), D + k)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    mask = cup.make_missing_mask(mat.shape, detect, detect, max_dist=D, sym_upper=True)
    kw = dict(max_dist=D, sym_upper=True, full=True, missing_tol=tol, pval=True)
    s = Session()
    s.upload(mat, kernel, missing_mask=mask, **kw)
    st = s.run()
    r, p = s.download()
    assert st["n_windows"] >= synthetic.n_windows(n, D)
    assert r.shape == (n, n) and r.nnz == st["nnz"] and abs(r.data).max() <= 1.0
    assert r.nnz > 0.9 * synthetic.n_windows(n, D)
    # the one-shot host call (slab-pipelined for a map of this size: uploads, kernels and
    # downloads overlap) returns exactly what the resident session does
    from chromosight_b200.utils import detection as cud
    r_h, p_h = cud.normxcorr2(mat, kernel, missing_mask=mask, **kw)
    for a_, b_ in ((r_h, r), (p_h, p)):
        assert np.array_equal(a_.indptr, b_.indptr) and np.array_equal(a_.indices, b_.indices)
        assert np.array_equal(a_.data, b_.data)
    del r_h, p_h
    # the same call with page-locked input arrays (direct DMA instead of staging)
    from chromosight_b200 import _cuda
    r_h, p_h = cud.normxcorr2(_cuda.pin_sparse(mat), kernel, missing_mask=_cuda.pin_sparse(mask), **kw)
    for a_, b_ in ((r_h, r), (p_h, p)):
        assert np.array_equal(a_.indptr, b_.indptr) and np.array_equal(a_.indices, b_.indices)
        assert np.array_equal(a_.data, b_.data)
    del r_h, p_h
    # candidates and foci: scores at / near the threshold are recomputed in float64 from the
    # CSR first (exact_refine), so the candidate set does not depend on float32 rounding
    cand, nc = s.candidates(pearson, 0, D)
    rec = records_to_numpy(cand, nc)
    foci = s.foci(pearson, 0, D, min_size=2)
    r2, p2 = s.download()                     # re-compacted from the refined image
    assert abs(r2 - r).max() <= SCORE_TOL
    rt = cup.diag_trim(r2, D).tocoo()
    sel = (rt.data >= pearson)
    assert nc == int(sel.sum())               # det:417-421 on the downloaded map
    key = np.sort(rec["row"].astype(np.int64) * n + rec["col"])
    assert np.array_equal(key, np.sort(rt.row[sel].astype(np.int64) * n + rt.col[sel]))
    foci_key = set((foci["row"].astype(np.int64) * n + foci["col"]).tolist())
    assert len(foci_key) == len(foci)
    # pattern_detector on the same map (det:177-345; upload and kernels pipelined slab by slab):
    # its patterns are these foci, minus the ones validate_patterns drops
    cfg = dict(getattr(presets, kname))
    cfg.update(pearson=pearson, max_perc_zero=100.0, max_perc_undetected=100 * tol, max_dist=D * 10000)
    table, wins = cud.pattern_detector(DummyMap(mat, D, (detect, detect), inter=False), cfg, kernel, full=True)
    tkey = set((table.bin1.values.astype(np.int64) * n + table.bin2.values).tolist())
    assert 0.5 * len(foci) < len(table) <= len(foci) and tkey <= foci_key
    assert wins.shape == (len(table), k, k)
    lut = {int(r_) * n + int(c_): float(s_) for r_, c_, s_ in zip(foci["row"], foci["col"], foci["score"])}
    assert max(abs(lut[int(b1) * n + int(b2)] - sc) for b1, b2, sc in
               zip(table.bin1.values, table.bin2.values, table.score.values)) <= 1e-7
    # oracle on crops
    rng = np.random.default_rng(5)
    W = D + 3 * k
    starts = [0, n - 400 - W] + list(rng.integers(100, n - 400 - W - 100, size=6))
    n_foci_checked = n_float_ties = 0
    for a0 in starts:
        a0 = int(a0)
        a1 = a0 + 400 + W
        sub = mat[a0:a1, a0:a1]
        msub = mask[a0:a1, a0:a1]
        r0, p0, nob = po.normxcorr2_dense(sub.toarray(), kernel, missing_mask=msub.toarray(),
                                          return_nobs=True, **kw)
        # interior: rows whose windows and mask geometry do not touch the crop's frame
        # (the first/last rows of the whole map are real edges and are compared too)
        lo = 0 if a0 == 0 else k
        hi = (a1 - a0) if a1 == n else (a1 - a0) - W
        got = r[a0 + lo:a0 + hi, a0:a1].toarray()
        got2 = r2[a0 + lo:a0 + hi, a0:a1].toarray()
        exp = r0[lo:hi, :].copy()
        gp = p[a0 + lo:a0 + hi, a0:a1].toarray()
        p0 = p0.copy()
        if a1 != n:
            # windows reaching beyond the crop's right edge are not comparable
            cols = np.arange(a1 - a0)[None, :] - (np.arange(lo, hi)[:, None])
            far = cols > D + 2 * k
            got[far] = 0
            got2[far] = 0
            exp[far] = 0
            gp[far] = 0
            p0[lo:hi][far] = 0
        _compare(got, gp, exp, p0[lo:hi], nob[lo:hi], strict=True)
        # refined scores: the candidates of the crop carry the float64 result itself
        # (row i of `exp` is row lo + i of the crop: map diagonals 0..D are its diagonals lo..D+lo)
        cpx = np.triu(np.tril(exp, D + lo), lo) >= pearson
        cpx &= exp != 0
        g2t = np.triu(np.tril(got2, D + lo), lo)
        assert np.array_equal((g2t >= pearson) & (g2t != 0), cpx)   # the candidate set, exactly
        # (all of them when they fit the refinement list of 8 M pixels)
        assert np.abs(got2[cpx] - exp[cpx].astype(np.float32)).max(initial=0) <= (1.2e-7 if nc < (1 << 23) else SCORE_TOL)
        # foci of the crop interior == pick_foci of the oracle (det:387-456): foci whose
        # pixels all lie at least 3 rows inside the compared rows are complete in the crop
        ex_trim = sp.coo_matrix(np.triu(np.tril(exp, D + lo), lo))
        coords0, lab0 = cud.pick_foci(ex_trim, pearson)
        exp_set = set()
        if coords0 is not None:
            lab0 = lab0.tocoo()
            for fid, (fy, fx) in zip(np.unique(lab0.data), coords0):
                sel_f = lab0.data == fid
                rows_f = lab0.row[sel_f]
                if rows_f.min() < 3 or rows_f.max() >= (hi - lo) - 3:
                    continue
                n_foci_checked += 1
                gkey = (int(fy) + a0 + lo) * n + (int(fx) + a0)
                if gkey in foci_key:
                    exp_set.add((int(fy) + a0 + lo, int(fx) + a0))
                    continue
                # The device keeps the scores as float32: two pixels of one focus whose float64
                # scores round to the same float32 tie, and the first one in row-major order wins
                # (np.argmax on the float64 map picks the larger).  Only such a tie may move a
                # focus' maximum, and only inside the focus.
                members = {(int(r_) + a0 + lo) * n + int(c_) + a0: (int(r_), int(c_))
                           for r_, c_ in zip(lab0.row[sel_f], lab0.col[sel_f])}
                mine = [members[key_] for key_ in members if key_ in foci_key]
                assert len(mine) == 1, (a0, fy, fx)
                assert np.float32(exp[mine[0]]) == np.float32(exp[fy, fx]), (a0, fy, fx, mine)
                n_float_ties += 1
                exp_set.add((mine[0][0] + a0 + lo, mine[0][1] + a0))
        # and the device foci centred well inside the crop are the oracle's
        inside = [(fr, fc) for fr, fc in zip(foci["row"], foci["col"])
                  if a0 + lo + 6 <= fr < a0 + hi - 6]
        if inside:
            if coords0 is not None:
                exp_set |= {(int(y) + a0 + lo, int(x) + a0) for y, x in coords0}
            for fr, fc in inside:
                assert (int(fr), int(fc)) in exp_set, (a0, fr, fc)
    # float32 ties are a corner of huge foci (borders at 0.15: a tenth of all pixels are
    # candidates); the loops configuration of the metric has none
    assert n_float_ties <= (0 if kname == "loops" else max(3, n_foci_checked // 200)), n_float_ties
    assert n_foci_checked > 0


from conftest import DummyMap, detector_case_names, load_detector_case  # noqa: E402


@pytest.mark.parametrize("name", detector_case_names())
def test_pattern_detector_golden(name):
    """The callers of the hot path: pattern_detector with the device-side
    thresholding, window gather / validation and score lookups, against the reference's
    own tables.  Foci coordinates and windows must be identical; scores within 1e-5;
    log10 p-values within 1e-4 relative over the score range where foci live."""
    from chromosight_b200.utils import detection as cud
    cmap, meta, kernel, coords, exp = load_detector_case(name)
    table, windows = cud.pattern_detector(cmap, meta["config"], kernel, coords=coords, full=meta["full"])
    if exp is None:
        assert table is None and windows is None
        return
    assert list(table.columns) == ["bin1", "bin2", "score", "pvalue"]
    assert np.array_equal(table.bin1.to_numpy(), exp["bin1"])
    assert np.array_equal(table.bin2.to_numpy(), exp["bin2"])
    assert windows.shape == exp["windows"].shape
    assert np.array_equal(np.isnan(windows), np.isnan(exp["windows"]))
    assert np.allclose(windows, exp["windows"], rtol=1e-15, atol=0, equal_nan=True)
    sc = table.score.to_numpy()
    assert np.array_equal(np.isnan(sc), np.isnan(exp["score"]))
    assert np.nanmax(np.abs(sc - exp["score"]), initial=0) <= SCORE_TOL
    lp, lp0 = np.log10(table.pvalue.to_numpy()), np.log10(exp["pvalue"])
    sel = np.isfinite(lp0) & (lp0 != 0)
    assert np.array_equal(lp0 == 0, lp == 0)
    # log10 p is ill-conditioned in r as |r| -> 1 (d log10p / dr ~ n / (1 - r^2)): the 1e-5
    # score tolerance maps to a few 1e-4 relative there, so compare with 1e-3
    assert np.allclose(lp[sel], lp0[sel], rtol=1e-3, atol=1e-4)


def test_validate_patterns_standalone(presets):
    """validate_patterns as a function of explicit matrices (det:18-155) against the oracle."""
    from chromosight_b200.utils import detection as cud
    from oracle import detector_oracle as do
    rng = np.random.default_rng(9)
    n, m = 120, 90
    a = rng.poisson(3.0, size=(n, m)).astype(float) * (rng.random((n, m)) < 0.8)
    a[rng.random((n, m)) < 0.01] = np.nan           # stored NaNs (the sub-diagonals of det:300-310)
    conv = sp.random(n, m, density=0.3, random_state=3, format="csr")
    vr = np.flatnonzero(rng.random(n) > 0.05)
    vc = np.flatnonzero(rng.random(m) > 0.05)
    coords = np.c_[rng.integers(0, n, 200), rng.integers(0, m, 200)]
    kernel = presets.loops_small["kernels"][0][:, 1:6]          # 7 x 5
    mr = np.ones(n, bool); mr[vr] = False
    mc = np.ones(m, bool); mc[vc] = False
    w0, ok0, s0 = do.validate_patterns_dense(coords, a, conv.toarray(), mr, mc, kernel.shape, 0.4, 0.6)
    for drop in (True, False):
        tab, win = cud.validate_patterns(coords, sp.csr_matrix(a), conv, (vr, vc), kernel, drop=drop,
                                         zero_tol=0.4, missing_tol=0.6)
        if drop:
            assert np.array_equal(tab.bin1.to_numpy(), coords[ok0, 0])
            assert np.allclose(win, w0[ok0], rtol=0, atol=0, equal_nan=True)
            assert np.allclose(tab.score.to_numpy(), s0[ok0])
        else:
            assert len(tab) == 200 and np.allclose(win, w0, rtol=0, atol=0, equal_nan=True)
            assert np.allclose(tab.score.to_numpy(), s0, equal_nan=True)
    assert 0 < ok0.sum() < 200


@pytest.mark.gpu
@pytest.mark.parametrize("thr", [0.3, 0.15, 0.05])
def test_device_foci_match_host_pick_foci(thr, presets):
    """pick_foci (det:387-456) on the device (4-connected labelling by union-find, size filter,
    per-focus argmax) against the host mirror, which is pinned to the reference
    (test_foci_match_reference): same foci, same order, same local maxima."""
    import scipy.sparse as sp
    from chromosight_b200 import synthetic
    from chromosight_b200.session import Session, records_to_numpy
    from chromosight_b200.utils import detection as cud, preprocessing as cup
    kernel = np.asarray(presets.loops["kernels"][0], dtype=np.float64)
    n, D, k = 6000, 120, kernel.shape[0]
    raw, detect = synthetic.band_counts(n, D + k, seed=11, missing_frac=0.02, max_dist=D)
    mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
    mat = cup.diag_trim(mat.tocsr(), D + k)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    mask = cup.make_missing_mask(mat.shape, detect, detect, max_dist=D, sym_upper=True)
    sess = Session()
    try:
        sess.upload(mat, kernel, max_dist=D, sym_upper=True, full=True, missing_mask=mask,
                    missing_tol=0.5, pval=True)
        sess.run()
        foci = sess.foci(thr, 0, D, min_size=2, cap=64)        # small cap: exercises the regrow
        rec, nc = sess.candidates(thr, 0, D)
        cand = records_to_numpy(rec, nc)
    finally:
        sess.close()
    cmat = sp.coo_matrix((cand["score"].astype(np.float64), (cand["row"], cand["col"])), shape=mat.shape)
    coords, labelled = cud.pick_foci(cmat, thr)
    assert coords is not None and len(coords) > 20
    assert len(foci) == len(coords)
    assert np.array_equal(np.stack([foci["row"], foci["col"]], axis=1), coords)
    # sizes and first pixels agree with the labelled matrix of the host
    lab = labelled.tocsr()
    ids = np.asarray(lab[foci["first_row"], foci["first_col"]]).ravel()
    assert np.array_equal(ids, np.sort(ids)) and len(np.unique(ids)) == len(ids)
    sizes = np.bincount(labelled.data.astype(np.int64))[ids.astype(np.int64)]
    assert np.array_equal(sizes, foci["size"])


def test_enqueued_run_equals_synchronous_run(presets):
    """Session.run(wait=False) + candidates() + wait() (one host synchronisation per step, what
    bench.py times) gives the results of the synchronous sequence, run after run."""
    from chromosight_b200 import synthetic
    from chromosight_b200.session import Session, records_to_numpy
    from chromosight_b200.utils import preprocessing as cup
    kernel = presets.loops["kernels"][0]
    k, n, D, thr = kernel.shape[0], 6000, 120, 0.3
    raw, detect = synthetic.band_counts(n, D + k, seed=23, missing_frac=0.03, max_dist=D)
    mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
    mat = cup.diag_trim(mat.tocsr(), D + k)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    kw = dict(max_dist=D, sym_upper=True, full=True, missing_tol=0.5, pval=True)
    s = Session()
    s.upload(mat, kernel, mask_geometry=cup.missing_geometry(mat.shape, detect, detect, D, True), **kw)
    st0 = s.run()
    rec0, n0 = s.candidates(thr, 0, D)
    rec0 = np.sort(records_to_numpy(rec0, n0).copy(), order=["row", "col"])
    r0, p0 = s.download()
    for _ in range(3):
        assert s.run(wait=False) is None
        rec, nc = s.candidates(thr, 0, D)
        st = s.wait()
        assert st["nnz"] == st0["nnz"] and st["n_windows"] == st0["n_windows"] and st["ms_pearson"] > 0
        rec = np.sort(records_to_numpy(rec, nc).copy(), order=["row", "col"])
        assert nc == n0 and np.array_equal(rec, rec0)
    s.run(wait=False)
    r1, p1 = s.download()          # download settles the enqueued run itself
    for a_, b_ in ((r0, r1), (p0, p1)):
        assert np.array_equal(a_.indptr, b_.indptr) and np.array_equal(a_.indices, b_.indices)
    # (r0 was downloaded after the refinement at thr: compare with a refined download)
    s.candidates(thr, 0, D)
    r2, _ = s.download()
    assert np.array_equal(r0.data, r2.data)
    s.close()


def test_row_slabs_match_single_run(presets):
    """SURVEY 8e, one chromosome over several GPUs: the row-slab path (rowslab.py; here the
    slabs of a 3-rank plan run one after the other on one device) yields exactly the candidate
    pixels and foci of a single run over the whole map.  Exact because scores at / near the
    threshold are recomputed in float64 from the CSR on both sides."""
    from chromosight_b200 import rowslab, synthetic
    from chromosight_b200.session import Session, records_to_numpy
    from chromosight_b200.utils import detection as cud, preprocessing as cup
    kernel = presets.loops["kernels"][0]
    k, n, D, thr, world = kernel.shape[0], 30_000, 200, 0.3, 3
    raw, detect = synthetic.band_counts(n, D + k, seed=17, missing_frac=0.02, max_dist=D)
    kw = dict(max_dist=D, sym_upper=True, full=True, missing_tol=0.5, pval=True)

    def candidates_of(mat, det, out_rows=None):
        s = Session()
        try:
            s.upload(mat, kernel, mask_geometry=cup.missing_geometry(mat.shape, det, det, D, True),
                     out_rows=out_rows, **kw)
            s.run(compact=False)
            rec, nc = s.candidates(thr, 0, D)
            return records_to_numpy(rec, nc).copy()
        finally:
            s.close()

    # single run
    mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
    mat = cup.diag_trim(mat.tocsr(), D + k)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    whole = rowslab.merge_sorted([candidates_of(mat, detect)])
    # slabs: partial law sums of the owned rows, summed like the all-reduce does
    plans = rowslab.slab_plan(n, world, k, D)
    sums = [rowslab.law_sums(rowslab.owned_rows(raw, p[0], p[1]), detect, D + k) for p in plans]
    tot_s = sum(s_.cpu().numpy() for s_, _ in sums)
    tot_c = sum(c_.cpu().numpy() for _, c_ in sums)
    law = rowslab.law_from_sums(tot_s, tot_c, n)
    assert np.allclose(law[: D + k + 1], np.nan_to_num(cup.distance_law(raw, detect, D + k, smooth=False))[: D + k + 1],
                       rtol=1e-12)
    parts = []
    for p in plans:
        sub, det = rowslab.slab_inputs(raw, detect, law, p, D, k)
        # every second slab scores its owned rows only (what bench.py --scaling strong does),
        # the others the whole sub-matrix: the owned candidates are the same
        rows = (p[0] - p[2], p[1] - p[2]) if len(parts) % 2 == 0 else None
        rec = candidates_of(sub, det, rows)
        if rows is not None:
            assert ((rec["row"] >= rows[0]) & (rec["row"] < rows[1])).all()
        parts.append(rowslab.owned_candidates(rec, p))
    merged = rowslab.merge_sorted(parts)
    assert len(whole) > 100
    assert np.array_equal(merged["row"], whole["row"]) and np.array_equal(merged["col"], whole["col"])
    assert np.abs(merged["score"] - whole["score"]).max() <= 2e-7
    c_m = rowslab.foci_of_candidates(merged, (n, n), thr)
    c_w = rowslab.foci_of_candidates(whole, (n, n), thr)
    assert np.array_equal(c_m, c_w)
