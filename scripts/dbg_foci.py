#!/usr/bin/env python3
"""Debug: device foci vs pick_foci of the oracle on the last crop of the 9x9 borders map."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, scipy.sparse as sp
from chromosight_b200 import kernels, synthetic
from chromosight_b200.session import Session, records_to_numpy
from chromosight_b200.utils import detection as cud, preprocessing as cup
from oracle import pearson_oracle as po
kernel = cup.resize_kernel(kernels.borders["kernels"][0], factor=9 / 17)
k = kernel.shape[0]; n, D, pearson, tol = 200_000, 200, 0.15, 0.75
raw, detect = synthetic.band_counts(n, D + k, seed=11, missing_frac=0.02, max_dist=D)
mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
mat = cup.diag_trim(mat.tocsr(), D + k); mat.data[np.isnan(mat.data)] = 0; mat.eliminate_zeros()
mask = cup.make_missing_mask(mat.shape, detect, detect, max_dist=D, sym_upper=True)
kw = dict(max_dist=D, sym_upper=True, full=True, missing_tol=tol, pval=True)
s = Session(); s.upload(mat, kernel, missing_mask=mask, **kw); s.run()
cand, nc = s.candidates(pearson, 0, D); rec = records_to_numpy(cand, nc)
foci = s.foci(pearson, 0, D, min_size=2)
print("candidates", nc, "foci", len(foci))
r2, _ = s.download()
W = D + 3 * k; a0 = n - 400 - W; a1 = n
sub = mat[a0:a1, a0:a1]; msub = mask[a0:a1, a0:a1]
r0, _ = po.normxcorr2_dense(sub.toarray(), kernel, missing_mask=msub.toarray(), **{**kw, "pval": False})
lo, hi = k, a1 - a0
exp = np.triu(np.tril(r0[lo:hi], D + 0)) if False else r0[lo:hi]
ex_trim = np.zeros_like(exp)
rr, cc = np.indices(exp.shape); dd = cc - (rr + lo)
ex_trim[(dd >= 0) & (dd <= D)] = exp[(dd >= 0) & (dd <= D)]
coords0, lab0 = cud.pick_foci(sp.coo_matrix(ex_trim), pearson)
lab0 = lab0.tocoo()
fk = {(int(a), int(b)): i for i, (a, b) in enumerate(zip(foci["row"], foci["col"]))}
# device labelling on the host mirror from the downloaded refined map
g = r2[a0 + lo:a0 + hi, a0:a1].toarray()
gt = np.zeros_like(g); gt[(dd >= 0) & (dd <= D)] = g[(dd >= 0) & (dd <= D)]
print("candidate sets equal:", np.array_equal((gt >= pearson) & (gt != 0), (ex_trim >= pearson) & (ex_trim != 0)))
bad = 0
for fid, (fy, fx) in zip(np.unique(lab0.data), coords0):
    sel = lab0.data == fid
    rows_f = lab0.row[sel]
    if rows_f.min() < 3: continue
    key = (int(fy) + a0 + lo, int(fx) + a0)
    if key in fk: continue
    members = set(zip((lab0.row[sel] + a0 + lo).tolist(), (lab0.col[sel] + a0).tolist()))
    mine = [m for m in members if m in fk]
    bad += 1
    if bad <= 5:
        print("oracle focus", key, "size", int(sel.sum()), "rows", rows_f.min(), rows_f.max(), "device foci inside:", mine,
              "oracle max", ex_trim[fy, fx], "device score there", gt[fy, fx])
        # which device focus contains this pixel? brute force: device foci whose first pixel..
        near = [(int(a), int(b), float(sc), int(sz)) for a, b, sc, sz in zip(foci["row"], foci["col"], foci["score"], foci["size"])
                if abs(int(a) - key[0]) < 40 and abs(int(b) - key[1]) < 60]
        print("  device foci nearby:", near[:8])
print("mismatching oracle foci:", bad)
