"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def check_summary(g, prefix, t, rtol, atol=0.0, what=""):
    """Compare tensor ``t`` with a (norm, sum, sampled entries) fixture written by make_golden.summarize."""
    t = t.detach().double().cpu().reshape(-1)
    norm = float(g[prefix + "/norm"])
    idx = torch.from_numpy(g[prefix + "/idx"])
    val = torch.from_numpy(g[prefix + "/val"])
    scale = max(norm / max(np.sqrt(t.numel()), 1.0), 1e-30)          # rms magnitude of the tensor
    err_n = abs(t.norm().item() - norm)
    assert err_n <= rtol * norm + atol, "%s %s: norm %g vs golden %g" % (what, prefix, t.norm().item(), norm)
    err = (t[idx] - val).abs().max().item()
    assert err <= rtol * max(scale, val.abs().max().item()) + atol, \
        "%s %s: sampled max abs err %g (rms scale %g)" % (what, prefix, err, scale)


def rel_l2(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
