import glob
import json
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_coo(z, prefix, dtype=None):
    shape = tuple(int(v) for v in z[prefix + "_shape"])
    val = z[prefix + "_val"]
    if dtype is not None:
        val = val.astype(dtype)
    return sp.coo_matrix((val, (z[prefix + "_row"], z[prefix + "_col"])), shape=shape)


def golden_case_names():
    return sorted(
        os.path.basename(p)[len("case_"):-len(".npz")]
        for p in glob.glob(os.path.join(GOLDEN, "case_*.npz"))
    )


def load_case(name):
    """One normxcorr2 fixture: (signal csr, kernel, kwargs dict, corr dense, pval dense)."""
    z = np.load(os.path.join(GOLDEN, f"case_{name}.npz"))
    kw = json.loads(str(z["kwargs"]))
    signal = load_coo(z, "signal").tocsr()
    dense = kw.pop("dense")
    has_mask = kw.pop("has_mask")
    if has_mask:
        kw["missing_mask"] = load_coo(z, "mask", bool).tocsr()
    corr = load_coo(z, "corr")
    corr_d = corr.toarray()
    pval_d = None
    if "pval_at_corr" in z.files:
        pval_d = np.zeros(corr.shape)
        pval_d[corr.row, corr.col] = z["pval_at_corr"]
    return signal, z["kernel"], kw, dense, corr_d, pval_d


@pytest.fixture(scope="session")
def presets():
    from chromosight_b200 import kernels
    return kernels


def detector_case_names():
    return sorted(
        os.path.basename(p)[len("detector_"):-len(".npz")]
        for p in glob.glob(os.path.join(GOLDEN, "detector_*.npz"))
    )


class DummyMap:
    """Stand-in for contacts_map.ContactMap with the attributes pattern_detector reads
    (the reference's tests use the same double, tests/test_detection.py:88-100)."""

    def __init__(self, matrix, max_dist=None, detectable_bins=None, inter=False, name="dummy"):
        self.matrix = matrix
        self.max_dist = max_dist
        self.detectable_bins = detectable_bins
        self.inter = inter
        self.name = name


def load_detector_case(name):
    z = np.load(os.path.join(GOLDEN, f"detector_{name}.npz"))
    meta = json.loads(str(z["meta"]))
    cmap = DummyMap(load_coo(z, "matrix").tocsr(), meta["max_dist"], (z["detect_rows"], z["detect_cols"]),
                    meta["inter"])
    coords = z["coords_in"] if "coords_in" in z.files else None
    exp = None
    if not meta["none"]:
        exp = {k: z[k] for k in ("bin1", "bin2", "pvalue", "score", "windows")}
    return cmap, meta, z["kernel"], coords, exp
