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


# (core_type, compare_func, post_process, aggregate) combinations of Cost beyond mse/raw (tests/golden/make_golden.py)
COST_OPTION_CASES = [("latent", "mse", "raw", "sum"), ("latent", "mse", "raw", "min"), ("subimage", "psuedo_intersect", "raw", "sum"),
                     ("subimage", "psuedo_intersect", "raw", "min"), ("subimage", "mse", "negative_exp", "sum"),
                     ("final_recon", "mse", "negative_exp", "sum"), ("latent", "mse", "negative_exp", "min")]


def check_ranking(sorted_costs, order, ref_sorted, ref_order, exact, what=""):
    """Candidate ranking against the reference's: index-exact for the continuous distances; the thresholded pixel counts of
    'psuedo_intersect' may move by a pixel when a value sits on a threshold, so there the sorted costs must agree to one pixel
    (1 / (2 * 3 * 64 * 64) ~ 4e-5 of the ratio, x K goals) and the order wherever neighbouring costs are further apart."""
    sorted_costs, order = np.asarray(sorted_costs, dtype=np.float64), np.asarray(order)
    if exact:
        assert np.array_equal(order, ref_order), what
        assert np.abs(sorted_costs - ref_sorted).max() <= 2e-6 * max(1.0, np.abs(ref_sorted).max()), what
        return
    tol = 1e-3
    assert np.abs(sorted_costs - ref_sorted).max() <= tol, what
    gaps = np.diff(ref_sorted)
    firm = np.concatenate([[True], gaps > 2 * tol]) & np.concatenate([gaps > 2 * tol, [True]])
    assert np.array_equal(order[firm], np.asarray(ref_order)[firm]), what
