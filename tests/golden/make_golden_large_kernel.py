#!/usr/bin/env python3
"""Golden normxcorr2 cases for the widest kernels, produced by the UNMODIFIED reference:

    python tests/golden/make_golden_large_kernel.py

  prod_stripes31      31x31 0/1 kernel (stripes_right preset): the widest kernel of the tiled
                      CUDA kernel, and mask-kernel sums that are exactly 0;
  prod_centromeres81  81x81 kernel (centromeres preset, max_dist 0 -> scan 1 diagonal): beyond
                      the tiled kernel, served by the one-warp-per-window kernel;
  valid_random45      45x45 random kernel, valid mode, no mask.
"""
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import chromosight.kernels as ck  # noqa: E402
from make_golden import production_case, save_normxcorr2_case  # noqa: E402


def main():
    stripes = np.array(ck.stripes_right["kernels"][0])
    centro = np.array(ck.centromeres["kernels"][0])
    _, _, mat, kw = production_case(260, 40, stripes, 21, missing_tol=0.5)
    save_normxcorr2_case("prod_stripes31", mat, stripes, **kw)
    _, _, mat, kw = production_case(330, 1, centro, 22, missing_tol=0.5)
    save_normxcorr2_case("prod_centromeres81", mat, centro, **kw)
    rng = np.random.default_rng(23)
    sig = sp.random(160, 190, density=0.4, random_state=5, format="csr")
    save_normxcorr2_case("valid_random45", sig, rng.random((45, 45)) + 0.1, pval=True)




def detector_case():
    """pattern_detector with the 81 x 81 centromere preset (1-D pattern: max_dist 0), lowered
    threshold so that the synthetic map yields patterns."""
    import make_golden_detector as mgd
    cfg = dict(ck.centromeres)
    cfg["pearson"] = 0.02
    cfg["max_perc_zero"] = 95          # the lower triangle of a window on the diagonal is empty
    cfg["max_perc_undetected"] = 95    # and its 81 sub-diagonals count as missing (det:300-310)
    mat, det = mgd.dense_intra_map(420, 1, 81, seed=44, n_loops=0)
    mgd.run_case("detect_centromeres81", mgd.DummyMap(mat, 1, (det, det)), cfg, cfg["kernels"][0])


if __name__ == "__main__":
    main()
    detector_case()
