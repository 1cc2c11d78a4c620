"""Preset pattern kernels, exposed like ``chromosight.kernels``
(chromosight/kernels/__init__.py:21-44): one module attribute per preset, each a
dict with the preset's parameters and a ``kernels`` list of 2-D float64 arrays.

The matrices are data fixtures extracted from the reference by
``tests/golden/make_golden.py`` into ``presets.npz``.
"""
import json as _json
import os as _os

import numpy as _np

_z = _np.load(_os.path.join(_os.path.dirname(__file__), "presets.npz"))
_meta = _json.loads(str(_z["meta"]))
for _name, _cfg in _meta.items():
    _cfg = dict(_cfg)
    _n = _cfg.pop("n_kernels")
    _cfg["kernels"] = [_np.array(_z[f"{_name}_{_i}"]) for _i in range(_n)]
    globals()[_name] = _cfg
names = sorted(_meta)
