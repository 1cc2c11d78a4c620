#!/usr/bin/env python3
"""Generate the golden fixtures of tests/golden/ and the kernel presets of
chromosight_b200/kernels/ by running the UNMODIFIED reference
(koszullab/chromosight, mounted read-only at /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Each ``case_*.npz`` holds the inputs of one ``normxcorr2`` / ``xcorr2`` /
``detrend`` call (COO triplets, kernel, JSON-encoded keyword arguments) and the
reference's outputs (COO triplets, float64).  The fixtures are what pins
oracle/ (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_parity.py).
"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

import chromosight.kernels as ck  # noqa: E402
import chromosight.utils.detection as cud  # noqa: E402
import chromosight.utils.preprocessing as cup  # noqa: E402
from scipy.stats import multivariate_normal  # noqa: E402

from chromosight_b200 import synthetic  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def coo_fields(prefix, mat):
    mat = sp.coo_matrix(mat)
    return {
        prefix + "_row": mat.row.astype(np.int32),
        prefix + "_col": mat.col.astype(np.int32),
        prefix + "_val": np.asarray(mat.data),
        prefix + "_shape": np.array(mat.shape, dtype=np.int64),
    }


def save_normxcorr2_case(name, signal, kernel, dense=False, **kw):
    """Run the reference and store inputs + outputs."""
    args = dict(kw)
    mask = args.get("missing_mask")
    sig_in = signal.toarray() if dense else signal.tocsr()
    corr, pval = cud.normxcorr2(sig_in, np.array(kernel), **args)
    fields = {}
    fields.update(coo_fields("signal", signal))
    fields["kernel"] = np.asarray(kernel, dtype=np.float64)
    if mask is not None:
        fields.update(coo_fields("mask", mask))
    jkw = {k: v for k, v in args.items() if k != "missing_mask"}
    jkw["dense"] = dense
    jkw["has_mask"] = mask is not None
    fields["kwargs"] = np.array(json.dumps(jkw))
    fields.update(coo_fields("corr", sp.coo_matrix(corr)))
    if pval is not None:
        # -inf / nan are legal p-values: store them densely aligned to corr's pattern
        c = sp.coo_matrix(corr)
        p = pval if dense else pval.toarray()
        fields["pval_at_corr"] = np.asarray(p)[c.row, c.col].astype(np.float64)
    np.savez_compressed(os.path.join(OUT, f"case_{name}.npz"), **fields)
    nnz = sp.coo_matrix(corr).nnz
    print(f"  {name}: shape={signal.shape} nnz_out={nnz}")


def gauss_mat(meanx, meany, std, shape=(100, 100)):
    """2-D Gaussian bump on a [-10, 10]^2 grid (the generator used by the
    reference's tests/test_detection.py:18-38)."""
    k = multivariate_normal(mean=(meanx, meany), cov=np.eye(2) * std)
    x = np.linspace(-10, 10, shape[0])
    y = np.linspace(-10, 10, shape[1])
    xx, yy = np.meshgrid(x, y)
    return sp.coo_matrix(k.pdf(np.c_[xx.ravel(), yy.ravel()]).reshape(shape))


def production_case(n, D, kernel, seed, missing_frac=0.03, missing_tol=0.5, tsvd=None):
    """What pattern_detector issues (det:242-263) on a synthetic intra map."""
    k = kernel.shape[0]
    raw, detect = synthetic.band_counts(n, D + k, seed=seed, missing_frac=missing_frac, max_dist=D)
    mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
    mat = cup.diag_trim(mat.tocsr(), D + k)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    mask = cup.make_missing_mask(mat.shape, detect, detect, max_dist=D, sym_upper=True)
    kw = dict(max_dist=D, sym_upper=True, full=True, missing_mask=mask,
              missing_tol=missing_tol, tsvd=tsvd, pval=True)
    return raw, detect, mat, kw


def main():
    os.makedirs(OUT, exist_ok=True)
    # ---- kernel presets (data fixtures of chromosight/kernels/) -------------
    presets = {}
    meta = {}
    for name in ("loops", "loops_small", "borders", "hairpins", "centromeres",
                 "stripes_left", "stripes_right"):
        cfg = getattr(ck, name)
        meta[name] = {k: v for k, v in cfg.items() if k != "kernels"}
        meta[name]["n_kernels"] = len(cfg["kernels"])
        for i, km in enumerate(cfg["kernels"]):
            presets[f"{name}_{i}"] = np.asarray(km, dtype=np.float64)
    presets["meta"] = np.array(json.dumps(meta))
    kdir = os.path.join(REPO, "chromosight_b200", "kernels")
    os.makedirs(kdir, exist_ok=True)
    np.savez_compressed(os.path.join(kdir, "presets.npz"), **presets)
    print("kernel presets written")

    loops = np.array(ck.loops["kernels"][0])
    loops_small = np.array(ck.loops_small["kernels"][0])
    borders = [np.array(k) for k in ck.borders["kernels"]]
    hairpin = np.array(ck.hairpins["kernels"][0])
    gk = gauss_mat(0, 0, 5, shape=(7, 7)).todense()
    gauss_kernel = np.asarray(gk + gk.T - np.diag(np.diag(gk)))

    print("normxcorr2 cases")
    # ---- reference test fixtures: Gaussians, valid mode, no mask ------------
    for i, (mx, my, sd) in enumerate([(-1.5, -1.0, 0.3), (-0.5, 1.0, 1.2), (0.5, 1.0, 2.7)]):
        g = gauss_mat(mx, my, sd).tocsr()
        save_normxcorr2_case(f"gauss{i}_sparse", g, gauss_kernel,
                             max_dist=None, sym_upper=False, pval=True)
        save_normxcorr2_case(f"gauss{i}_dense", g, gauss_kernel, dense=True,
                             max_dist=None, sym_upper=False, pval=True)
    # ---- built-in kernels planted at (60, 80), det tests :340-364 -----------
    for nm, kern in (("loops", loops), ("borders0", borders[0]), ("hairpin", hairpin)):
        km, kn = kern.shape
        sig = np.zeros((100, 100))
        sig[60 - km // 2:60 + km // 2 + 1, 80 - kn // 2:80 + kn // 2 + 1] = kern
        save_normxcorr2_case(f"planted_{nm}", sp.csr_matrix(np.triu(sig)), kern,
                             max_dist=None, sym_upper=False, pval=True)
    # ---- masked / full / sym_upper without max_dist (test_missing_corr) -----
    for nm, kern in (("loops", loops), ("hairpin", hairpin)):
        mat = sp.csr_matrix(cup.resize_kernel(kern, factor=10)).tolil()
        mask = sp.lil_matrix(mat.shape, dtype=bool)
        cm = mat.shape[0] // 2
        rows = np.array([-2, 1, 2]) + cm
        mat[rows, :] = 0.0
        mask[rows, :] = True
        mat = sp.triu(mat.tocsr()).tocsr()
        save_normxcorr2_case(f"missing_rows_{nm}", mat, kern, missing_mask=mask.tocsr(),
                             sym_upper=True, full=True, pval=True)
    # ---- production call on synthetic intra maps ----------------------------
    for nm, n, D, kern, seed, tol, tsvd in (
        ("prod_loops17", 300, 40, loops, 1, 0.5, None),
        ("prod_loops7", 220, 30, loops_small, 2, 0.5, None),
        ("prod_borders9", 260, 30, cup.resize_kernel(borders[1], factor=9 / 17), 3, 0.75, None),
        ("prod_hairpin15_d1", 200, 1, hairpin, 4, 0.75, None),
        ("prod_loops17_tsvd", 200, 30, loops, 5, 0.5, 0.999),
        ("prod_tiny_n", 40, 60, loops, 6, 0.5, None),
    ):
        _, _, mat, kw = production_case(n, D, kern, seed, missing_tol=tol, tsvd=tsvd)
        save_normxcorr2_case(nm, mat, kern, **kw)
    # full mode, no mask
    _, _, mat, kw = production_case(200, 30, loops_small, 7)
    kw["missing_mask"] = None
    save_normxcorr2_case("full_nomask", mat, loops_small, **kw)
    # valid mode with a (non framed) mask
    _, _, mat, kw = production_case(150, 30, loops_small, 8)
    kw["full"] = False
    save_normxcorr2_case("valid_mask", mat, loops_small, **kw)
    # ---- inter-chromosomal (rectangular, whole rows/cols missing) -----------
    imat, (vr, vc) = synthetic.inter_counts(120, 90, seed=9, density=0.3, missing_frac=0.05)
    imask = cup.make_missing_mask(imat.shape, vr, vc, max_dist=None, sym_upper=False)
    save_normxcorr2_case("inter_rect", imat, loops_small, max_dist=None, sym_upper=False,
                         full=True, missing_mask=imask, missing_tol=0.75, pval=True)

    # ---- xcorr2 -------------------------------------------------------------
    print("xcorr2 cases")
    g = gauss_mat(-0.5, 1.0, 1.2).tocsr()
    _, _, mat, _ = production_case(150, 30, loops_small, 10)
    xc = {}
    for nm, sig, kern, kw in (
        ("gauss", g, gauss_kernel, {}),
        ("band", mat, loops_small, {}),
        ("band_const", mat, np.ones((5, 5)) * 0.25, {}),
        ("band_tsvd", mat, loops, {"tsvd": 0.999}),
        ("band_rect", mat, loops[4:13, :], {}),
    ):
        out = cud.xcorr2(sig, kern, **kw)
        xc.update(coo_fields(f"{nm}_signal", sig))
        xc[f"{nm}_kernel"] = np.asarray(kern, dtype=np.float64)
        xc[f"{nm}_kwargs"] = np.array(json.dumps(kw))
        xc.update(coo_fields(f"{nm}_out", out))
    np.savez_compressed(os.path.join(OUT, "xcorr2_cases.npz"), **xc)

    # ---- detrend / distance_law / masks -------------------------------------
    print("preprocessing cases")
    pp = {}
    for i, (n, D, seed) in enumerate(((300, 57, 11), (120, 200, 12))):
        raw, detect = synthetic.band_counts(n, D, seed=seed, missing_frac=0.04)
        full_sym = (raw + sp.triu(raw, 1).T).tocsr()
        for tag, m in (("upper", raw), ("sym", full_sym)):
            law = cup.distance_law(m.tocsr(), detectable_bins=detect, max_dist=D, smooth=False)
            det = cup.detrend(m, detectable_bins=detect, max_dist=D, max_val=10)
            det_nomax = cup.detrend(m, detectable_bins=detect, max_dist=D, max_val=None)
            pp.update(coo_fields(f"d{i}_{tag}_raw", m))
            pp[f"d{i}_{tag}_detect"] = detect.astype(np.int64)
            pp[f"d{i}_{tag}_max_dist"] = np.array(D)
            pp[f"d{i}_{tag}_law"] = law
            pp.update(coo_fields(f"d{i}_{tag}_out", det))
            pp.update(coo_fields(f"d{i}_{tag}_out_nomax", det_nomax))
    # masks: make + frame for intra/inter
    vr = np.array([0, 1, 2, 4, 5, 7, 8, 9, 11])
    m1 = cup.make_missing_mask((12, 12), vr, vr, max_dist=3, sym_upper=True)
    f1 = cup.frame_missing_mask(m1, (5, 5), sym_upper=True, max_dist=3)
    m2 = cup.make_missing_mask((12, 9), vr, np.array([0, 2, 3, 4, 6, 8]), sym_upper=False)
    f2 = cup.frame_missing_mask(m2, (5, 3), sym_upper=False, max_dist=None)
    m3 = cup.make_missing_mask((12, 12), vr, vr, max_dist=None, sym_upper=True)
    f3 = cup.frame_missing_mask(m3, (3, 5), sym_upper=True, max_dist=None)
    f4 = cup.frame_missing_mask(m1, (3, 7), sym_upper=True, max_dist=2)
    for nm, m in (("m1", m1), ("f1", f1), ("m2", m2), ("f2", f2), ("m3", m3), ("f3", f3), ("f4", f4)):
        pp["mask_" + nm] = m.toarray()
    pp["mask_valid_rows"] = vr
    # ztransform / diag_trim
    zt = cup.ztransform(raw.tocoo())
    pp.update(coo_fields("zt_out", zt))
    pp.update(coo_fields("trim_out", cup.diag_trim(raw.tocsr(), 17)))
    # factorise_kernel
    u, v = cup.factorise_kernel(loops.copy(), prop_info=0.999)
    pp["fact_loops_uv"] = u @ v
    np.savez_compressed(os.path.join(OUT, "preproc_cases.npz"), **pp)

    # ---- foci picking (det:387-592) on a production correlation map ---------
    print("foci cases")
    fc = {}
    _, _, mat, kw = production_case(400, 40, loops, 13, missing_tol=0.5)
    corr, _ = cud.normxcorr2(mat, loops, **kw)
    corr = cup.diag_trim(corr.tocsr(), 40).tocoo()
    corr.eliminate_zeros()
    for thr in (0.2, 0.3):
        coords, lab = cud.pick_foci(corr.copy(), thr)
        fc[f"thr{int(thr*100)}_coords"] = np.zeros((0, 2), int) if coords is None else coords
        if lab is not None:
            fc.update(coo_fields(f"thr{int(thr*100)}_labels", lab))
    fc.update(coo_fields("corr", corr))
    np.savez_compressed(os.path.join(OUT, "foci_cases.npz"), **fc)
    print("done")


if __name__ == "__main__":
    main()
