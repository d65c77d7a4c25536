"""Loader for the golden vectors written by tests/golden/make_golden.py."""
import os
from collections import defaultdict

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_cache = {}


def load(fname):
    """-> {case: {group: {name: np.ndarray}}}"""
    if fname not in _cache:
        z = np.load(os.path.join(GOLDEN, fname), allow_pickle=False)
        out = defaultdict(lambda: defaultdict(dict))
        for key in z.files:
            case, group, name = key.split("|")
            out[case][group][name] = z[key]
        _cache[fname] = out
    return _cache[fname]


def tensors(d, device="cpu", dtype=None):
    out = {}
    for k, v in d.items():
        t = torch.from_numpy(np.ascontiguousarray(v))
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        out[k] = t.to(device)
    return out


def loss_weights(shape, i, device="cpu", dtype=torch.float32):
    """Same projection tensor as make_golden.loss_weights."""
    n = int(np.prod(shape))
    return torch.cos(torch.arange(n, dtype=torch.float64) * 0.37 + i).to(torch.float32) \
        .reshape(shape).to(device=device, dtype=dtype)


def meta_scalar(case, name):
    return case["meta"][name].item()


def kwargs_of(case):
    kw = {}
    for k, v in case["kw"].items():
        if v.dtype.kind in "US":
            kw[k] = str(v)
        elif v.dtype.kind == "b":
            kw[k] = bool(v)
        elif v.dtype.kind in "iu":
            kw[k] = int(v)
        else:
            kw[k] = float(v)
    return kw
