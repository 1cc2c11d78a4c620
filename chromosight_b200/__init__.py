"""chromosight_b200: the Pearson template-matching hot path of koszullab/chromosight on B200.

Drop-in modules (same names and signatures as the reference's):
    chromosight_b200.utils.detection      normxcorr2, xcorr2, pattern_detector, pick_foci, ...
    chromosight_b200.utils.preprocessing  detrend, distance_law, make_missing_mask, ...
    chromosight_b200.utils.stats          corr_to_pval, fdr_correction
    chromosight_b200.kernels              preset pattern kernels
Around them: session.Session (device-resident call), contacts_map / cool (genome model and
.cool reader without cooler), driver (rank-per-GPU detect / quantify), sharding.

Importing the package needs neither CUDA nor the shared library; every hot-path function raises
_lib.BackendError without them (there is no CPU fallback).
"""
__version__ = "0.1.0"
