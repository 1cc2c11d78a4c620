"""Loader of oracle/_ref: the unmodified reference modules of the hot path, copied from
/root/reference by oracle/make_ref.sh (git-ignored, shipped with the snapshot).

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import importlib
import os
import sys

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available():
    return os.path.exists(os.path.join(REF_DIR, "chromosight", "utils", "detection.py"))


def load():
    """(detection, preprocessing, stats) modules of the unmodified reference, or None."""
    if not available():
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        det = importlib.import_module("chromosight.utils.detection")
        pre = importlib.import_module("chromosight.utils.preprocessing")
        sta = importlib.import_module("chromosight.utils.stats")
    if not os.path.abspath(det.__file__).startswith(REF_DIR):
        return None  # another chromosight is installed and shadows the copy
    return det, pre, sta
