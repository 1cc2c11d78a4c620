"""Top-error windows of golden cases (dev tool, GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, scipy.sparse as sp
from conftest import load_case
from chromosight_b200.utils import detection as cud
from oracle import pearson_oracle as po
for name in sys.argv[1:]:
    signal, kernel, kw, dense, corr_ref, pval_ref = load_case(name)
    kw2 = dict(kw); kw2["pval"] = False
    r, _ = cud.normxcorr2(signal, kernel, **kw2)
    r = r.toarray()
    d = np.abs(r - corr_ref)
    A = signal.toarray(); n0, n1 = A.shape
    M = kw.get("missing_mask"); M = M.toarray() if M is not None else np.zeros_like(A, bool)
    kh, kw_ = kernel.shape[0] // 2, kernel.shape[1] // 2
    print(f"{name}: max err {d.max():.2e}; quantiles", np.quantile(d[corr_ref != 0], [0.5, 0.9, 0.99, 0.999]))
    for i in np.argsort(d.ravel())[::-1][:6]:
        y, x = divmod(i, n1)
        w = A[max(0, y-kh):y+kh+1, max(0, x-kw_):x+kw_+1]; mm = M[max(0, y-kh):y+kh+1, max(0, x-kw_):x+kw_+1]
        pres = w[~mm]
        print(f"  ({y},{x}) d={x-y} gpu={r[y,x]:+.7f} ref={corr_ref[y,x]:+.7f} err={d[y,x]:.2e} present: mean={pres.mean():.4f} std={pres.std():.5f} min={pres.min():.3f} max={pres.max():.3f} nmiss={int(mm.sum())} zero-filled std={w.std():.4f}")
