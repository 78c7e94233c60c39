"""Helpers for the GPU parity tests: build the op3_b200 model with oracle weights and inject the oracle's noise."""
import ctypes as C

import numpy as np
import torch

from oracle import op3_oracle as orc

VARIANT = dict(refinement_model_type="size_dependent_conv", decoder_model_type="reg", dynamics_model_type="reg_ac32",
               sto_repsize=64, det_repsize=64, extra_args=dict(beta=1e-2, deterministic_sampling=False))


def build_model(action_dim, K, seed=0, device="cuda"):
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    v = dict(VARIANT, K=K)
    m = create_model_v2(v, 64, 64, action_dim)
    p = orc.make_params(action_dim, seed=seed)
    m.load_state_dict(p, strict=True)
    return m.to(device), p


class NoiseInjector:
    """Replaces op3_b200.torch.pytorch_util.randn by a queue of pre-generated CPU tensors (like the golden generator
    does for the reference)."""

    def __init__(self, tensors, device="cuda"):
        self.q = [t.to(device) for t in tensors]

    def __enter__(self):
        from op3_b200.torch import pytorch_util as ptu
        self.ptu, self.orig = ptu, ptu.randn

        def fake(*size, **kw):
            t = self.q.pop(0)
            assert tuple(t.shape) == tuple(size), (t.shape, size)
            return t
        ptu.randn = fake
        return self

    def __exit__(self, *a):
        self.ptu.randn = self.orig


def rel_l2(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def max_abs(a, b):
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()


def compare_grads(ours, ref, total_norm, rtol, report=None, floor=1e-6):
    """Per-tensor relative L2 error with an absolute floor tied to the global gradient norm.  Returns worst."""
    worst = ("", 0.0)
    for k, g in ref.items():
        mine = ours[k]
        err = (mine.detach().double().cpu() - g.double()).norm().item()
        scale = max(g.double().norm().item(), floor * total_norm)
        r = err / scale
        if report is not None:
            report.append("%-60s rel %.3e  |ref| %.3e" % (k, r, g.norm().item()))
        if r > worst[1]:
            worst = (k, r)
    assert worst[1] <= rtol, "gradient mismatch: %s rel-L2 %.3e > %.1e" % (worst[0], worst[1], rtol)
    return worst
