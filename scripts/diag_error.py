"""Where does the score error of the CUDA path come from? (dev tool, GPU box)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
from chromosight_b200 import kernels, synthetic
from chromosight_b200.utils import detection as cud, preprocessing as cup
from oracle import pearson_oracle as po

kernel = kernels.loops["kernels"][0]
for (n, D, seed) in ((600, 50, 3), (700, 60, 21)):
    k = kernel.shape[0]
    raw, detect = synthetic.band_counts(n, D + k, seed=seed, missing_frac=0.03, max_dist=D)
    mat = cup.detrend(raw, detectable_bins=detect, max_dist=D + k, max_val=10)
    mat = cup.diag_trim(mat.tocsr(), D + k)
    mat.data[np.isnan(mat.data)] = 0
    mat.eliminate_zeros()
    mask = cup.make_missing_mask(mat.shape, detect, detect, max_dist=D, sym_upper=True)
    kw = dict(max_dist=D, sym_upper=True, full=True, missing_tol=0.5, pval=False)
    r, _ = cud.normxcorr2(mat, kernel, missing_mask=mask, **kw)
    r = r.toarray()
    r0, _ = po.normxcorr2_dense(mat.toarray(), kernel, missing_mask=mask.toarray(), **kw)
    m32 = mat.copy(); m32.data = m32.data.astype(np.float32).astype(np.float64)
    r32, _ = po.normxcorr2_dense(m32.toarray(), kernel, missing_mask=mask.toarray(), **kw)
    d = np.abs(r - r0)
    print(f"n={n}: max|gpu-ref|={d.max():.2e}  max|ref(f32 input)-ref|={np.abs(r32-r0).max():.2e}  max|gpu-ref(f32 input)|={np.abs(r-r32).max():.2e}")
    idx = np.argsort(d.ravel())[::-1][:8]
    A = mat.toarray(); M = mask.toarray()
    for i in idx:
        y, x = divmod(i, n)
        y0, y1, x0, x1 = max(0, y-8), min(n, y+9), max(0, x-8), min(n, x+9)
        w = A[y0:y1, x0:x1]; mm = M[y0:y1, x0:x1]
        pres = w[~mm]
        print(f"  ({y},{x}) d={x-y} gpu={r[y,x]:+.7f} ref={r0[y,x]:+.7f} ref32={r32[y,x]:+.7f} err={d[y,x]:.2e} win mean={pres.mean():.3f} std={pres.std():.3f} nmiss={int(mm.sum())} nz={int((pres!=0).sum())}/{pres.size}")
    q = np.quantile(d[r0 != 0], [0.5, 0.9, 0.99, 0.999])
    print("  quantiles of |err| over nonzero windows:", q)
