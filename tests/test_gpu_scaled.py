"""GPU: the scaled configuration of BASELINE.json (128x128 frames, more slots) through new registry entries, against the
CPU oracle at a size the oracle finishes in seconds (B=1).  Exercises D=128 tiling, K up to 14 (16-slot kernel variant)
and the Rd/K-generic paths."""
import numpy as np
import pytest
import torch

import op3_b200
from oracle import op3_oracle as orc
from gpu_util import NoiseInjector, rel_l2, max_abs, compare_grads, VARIANT

pytestmark = pytest.mark.gpu


def build_scaled(K, D, action_dim, seed=0):
    from op3_b200.torch.op3_modules import decoder_network, refinement_network
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    v = dict(VARIANT, K=K, decoder_model_type=decoder_network.register_image_size(D),
             refinement_model_type=refinement_network.register_image_size(D))
    m = create_model_v2(v, 64, 64, action_dim)
    p = orc.make_params(action_dim, seed=seed)
    m.load_state_dict(p, strict=True)
    return m.cuda(), p


@pytest.mark.parametrize("K,precision,D", [(14, "fp32", 128), (8, "tf32x3", 128), (8, "f16x3", 128), (3, "f16x3", 96),
                                           (3, "f16x3", 80)])
def test_scaled_128_matches_oracle(K, precision, D):
    """f16x3 beyond D = 88 / 72: the quint-tap and ring weight-gradient kernels walk the image in column strips (128 -> 2 x 64,
    96 -> 2 x 48; 80 is one item for the quint kernel and two strips of 40 for the ring kernel)."""
    B, T, sched = 1, 2, [0, 0, 1]
    op3_b200.set_precision(precision)
    try:
        m, p = build_scaled(K, D, 0)
        g = torch.Generator().manual_seed(3)
        frames = torch.randint(0, 256, (B, T, 3, D, D), generator=g, dtype=torch.uint8).float() / 255.0
        lw = np.arange(1, len(sched) + 1)
        with NoiseInjector(orc.make_noise(B, K, 64, len(sched) + 1)):
            out = m.forward(frames.cuda(), None, None, np.array(sched, dtype=np.float64), lw)
        out[3].backward()
        grads = {k: v.grad for k, v in m.named_parameters()}
    finally:
        op3_b200.set_precision("fp32")
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    o = orc.run_schedule(q, frames, None, sched, lw, orc.make_noise(B, K, 64, len(sched) + 1), K, D=D)
    o["total_loss"].backward()
    ograds = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in q.items()}
    tol = dict(loss=1e-5, masks=1e-5, grads=1e-4, floor=1e-6) if precision == "fp32" else dict(loss=2e-5, masks=2e-5, grads=2e-4, floor=1e-5)
    lrel = abs(out[3].item() - o["total_loss"].item()) / abs(o["total_loss"].item())
    merr = max_abs(out[1], o["masks_list"])
    print("scaled D=%d K=%d %s: loss rel %.2e masks %.2e" % (D, K, precision, lrel, merr))
    assert lrel <= tol["loss"] and merr <= tol["masks"]
    tn = torch.sqrt(sum((v.double() ** 2).sum() for v in ograds.values())).item()
    compare_grads(grads, ograds, tn, tol["grads"], floor=tol["floor"])
