"""The oracle (oracle/pearson_oracle.py) against outputs of the unmodified
reference stored in tests/golden (generator: tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_case_names, load_case, load_coo
from oracle import pearson_oracle as po


@pytest.mark.parametrize("name", golden_case_names())
def test_normxcorr2_oracle_matches_reference(name):
    signal, kernel, kw, dense, corr_ref, pval_ref = load_case(name)
    mask = kw.pop("missing_mask", None)
    pval = kw.pop("pval", False)
    r, p = po.normxcorr2_dense(
        signal.toarray(), kernel,
        missing_mask=None if mask is None else mask.toarray(), pval=pval, **kw)
    assert r.shape == corr_ref.shape
    # tolerance: fp64 round-off of two different summation orders.  Windows of
    # exactly constant signal have a variance that is pure round-off in the
    # reference (|r| ~ 1e-8 noise, e.g. the zoomed 0.5/1.5 hairpin template):
    # there only the magnitude is comparable.
    d = np.abs(r - corr_ref)
    sig = (np.abs(corr_ref) > 1e-6) | (np.abs(r) > 1e-6)
    assert d.max() < 1e-6
    assert d[sig].max() < 1e-8
    if pval_ref is not None:
        nz = (corr_ref != 0) & sig
        fin = nz & np.isfinite(pval_ref)
        assert np.array_equal(np.isneginf(p[nz]), np.isneginf(pval_ref[nz]))
        assert np.allclose(p[fin], pval_ref[fin], rtol=1e-6, atol=1e-7)


def test_xcorr2_oracle_matches_reference():
    import json
    z = np.load(os.path.join(GOLDEN, "xcorr2_cases.npz"))
    for nm in ("gauss", "band", "band_const", "band_tsvd", "band_rect"):
        sig = load_coo(z, f"{nm}_signal").toarray()
        kw = json.loads(str(z[f"{nm}_kwargs"]))
        out = po.xcorr2_dense(sig, z[f"{nm}_kernel"], **kw)
        ref = load_coo(z, f"{nm}_out").toarray()
        assert np.max(np.abs(out - ref)) < 1e-9, nm


def test_preprocessing_oracle_matches_reference():
    z = np.load(os.path.join(GOLDEN, "preproc_cases.npz"))
    for i in range(2):
        for tag in ("upper", "sym"):
            raw = load_coo(z, f"d{i}_{tag}_raw").toarray()
            det = z[f"d{i}_{tag}_detect"]
            D = int(z[f"d{i}_{tag}_max_dist"])
            law = po.distance_law_dense(raw, det, D)
            assert np.allclose(law, z[f"d{i}_{tag}_law"], rtol=1e-12, equal_nan=True)
            out = po.detrend_dense(raw, det, D, max_val=10)
            ref = load_coo(z, f"d{i}_{tag}_out").toarray()
            assert np.allclose(out, ref, rtol=1e-12, atol=0)
            out = po.detrend_dense(raw, det, D, max_val=None)
            ref = load_coo(z, f"d{i}_{tag}_out_nomax").toarray()
            assert np.allclose(out, ref, rtol=1e-12, atol=0)
    vr = z["mask_valid_rows"]
    m1 = po.make_missing_mask_dense((12, 12), vr, vr, max_dist=3, sym_upper=True)
    assert np.array_equal(m1, z["mask_m1"])
    assert np.array_equal(po.frame_missing_mask_dense(m1, (5, 5), True, 3), z["mask_f1"])
    m2 = po.make_missing_mask_dense((12, 9), vr, np.array([0, 2, 3, 4, 6, 8]), sym_upper=False)
    assert np.array_equal(m2, z["mask_m2"])
    assert np.array_equal(po.frame_missing_mask_dense(m2, (5, 3), False, None), z["mask_f2"])
    m3 = po.make_missing_mask_dense((12, 12), vr, vr, max_dist=None, sym_upper=True)
    assert np.array_equal(m3, z["mask_m3"])
    assert np.array_equal(po.frame_missing_mask_dense(m3, (3, 5), True, None), z["mask_f3"])
    assert np.array_equal(po.frame_missing_mask_dense(m1, (3, 7), True, 2), z["mask_f4"])
    assert np.allclose(po.truncate_kernel(__import__("chromosight_b200.kernels", fromlist=["x"]).loops["kernels"][0], 0.999),
                       z["fact_loops_uv"], atol=1e-12)


def test_reference_known_answers():
    """Known answers of the reference's own tests: distance law of the docstring
    example (pre:162-171, tests/test_preprocessing.py:202-213) and
    make_missing_mask's worked example (pre:578-585)."""
    m = np.ones((3, 3)) + np.array([1, 2, 3])
    assert np.allclose(po.distance_law_dense(m), [3.0, 3.5, 4.0])
    valid = np.array([0, 2, 4])
    exp = np.array([[0, 1, 0, 0, 0], [0, 1, 1, 0, 0], [0, 0, 0, 1, 0], [0, 0, 0, 1, 1], [0, 0, 0, 0, 0]], bool)
    assert np.array_equal(po.make_missing_mask_dense((5, 5), valid, valid, max_dist=1, sym_upper=True), exp)


from conftest import detector_case_names, load_detector_case  # noqa: E402


@pytest.mark.parametrize("name", detector_case_names())
def test_detector_oracle_matches_reference(name):
    """pattern_detector / validate_patterns restatement against the reference's own output
    (coordinates, windows, p-values; scores as the trimmed map read at the coordinates)."""
    from chromosight_b200.utils.detection import pick_foci      # host code, pinned by foci_cases.npz
    from oracle import detector_oracle as do
    cmap, meta, kernel, coords, exp = load_detector_case(name)
    out = do.pattern_detector_dense(cmap.matrix.toarray(), cmap.detectable_bins, cmap.max_dist, cmap.inter,
                                    meta["config"], kernel, pick_foci, coords=coords, full=meta["full"])
    if exp is None:
        assert out is None
        return
    b1, b2, score, pvalue, windows = out
    assert np.array_equal(b1, exp["bin1"]) and np.array_equal(b2, exp["bin2"])
    assert np.allclose(windows, exp["windows"], rtol=1e-12, atol=0, equal_nan=True)
    assert np.allclose(score, exp["score"], rtol=0, atol=1e-9, equal_nan=True)
    assert np.allclose(np.log10(pvalue), np.log10(exp["pvalue"]), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", golden_case_names())
def test_sparse_port_matches_reference(name):
    """oracle/sparse_port.py (the timed CPU port of bench.py) against the reference's outputs on
    every sparse fixture.  Stored round-off noise of exactly flat windows (|r| < 1e-6) may
    differ between the two summation orders, everything else agrees to 1e-8 (1e-14 typically;
    the near-flat windows of the zoomed hairpin template amplify fp64 round-off)."""
    from oracle import sparse_port as spt
    signal, kernel, kw, dense, corr_ref, pval_ref = load_case(name)
    if dense:
        pytest.skip("dense fixture: the port follows the sparse path (det:917-1131)")
    r, p = spt.normxcorr2_sparse(signal, kernel, **kw)
    r = r.toarray()
    d = np.abs(r - corr_ref)
    sig = (np.abs(corr_ref) > 1e-6) | (np.abs(r) > 1e-6)
    assert d.max() < 1e-6
    assert d[sig].max(initial=0) < 1e-8
    if pval_ref is not None:
        p = p.toarray()
        fin = sig & (corr_ref != 0) & np.isfinite(pval_ref)
        assert np.allclose(p[fin], pval_ref[fin], rtol=1e-6, atol=1e-7)


def test_oracle_matches_live_reference_when_present():
    """Where oracle/_ref (the unmodified reference, oracle/make_ref.sh) travelled with the
    repository, the dense oracle and the sparse port are checked against it live on a seeded
    production-style call -- not only through the stored fixtures."""
    from oracle import ref_loader, sparse_port as spt
    ref = ref_loader.load()
    if ref is None:
        pytest.skip("oracle/_ref absent (run oracle/make_ref.sh in the build container)")
    det, pre, _ = ref
    from chromosight_b200 import kernels, synthetic
    kernel = kernels.loops["kernels"][0]
    n, D, k = 400, 40, kernel.shape[0]
    raw, detect = synthetic.band_counts(n, D + k, seed=9, missing_frac=0.04, max_dist=D)
    mat = pre.detrend(raw.tocsr(), detectable_bins=detect, max_dist=D + k, max_val=10)
    mat = pre.diag_trim(mat.tocsr(), D + k)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    assert np.allclose(mat.toarray(), np.triu(np.tril(po.detrend_dense(raw.toarray(), detect, D + k, 10), D + k)),
                       rtol=1e-12)
    mask = pre.make_missing_mask(mat.shape, detect, detect, max_dist=D, sym_upper=True)
    kw = dict(max_dist=D, sym_upper=True, full=True, missing_tol=0.5, pval=True)
    r_ref, p_ref = det.normxcorr2(mat, kernel, missing_mask=mask, **kw)
    r0, p0 = po.normxcorr2_dense(mat.toarray(), kernel, missing_mask=mask.toarray(), **kw)
    assert np.abs(r_ref.toarray() - r0).max() < 1e-10
    r1, p1 = spt.normxcorr2_sparse(mat, kernel, missing_mask=mask, **kw)
    assert np.abs(r_ref.toarray() - r1.toarray()).max() < 1e-10
    nz = r0 != 0
    assert np.allclose(p_ref.toarray()[nz], p0[nz], rtol=1e-8, atol=1e-9)
