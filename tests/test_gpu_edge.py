"""GPU: edge shapes of the training step against the CPU oracle -- a single slot, a single sample, the smallest and
odd image sides the kernels accept (multiples of 8), the 16-slot maximum, zero loss weights (decodes without a loss,
trap T5) -- in the exact fp32 mode and in the default fp16-split tensor-core mode."""
import numpy as np
import pytest
import torch

import op3_b200
from oracle import op3_oracle as orc
from gpu_util import NoiseInjector, max_abs, compare_grads
from test_gpu_scaled import build_scaled

pytestmark = pytest.mark.gpu

SHAPES = [  # B, K, D, action_dim, schedule, loss weights
    (1, 1, 64, 0, [0, 1, 0], [1, 2, 3]),
    (3, 5, 32, 4, [0, 0, 1, 0], [0, 1, 0, 2]),
    (2, 16, 16, 0, [0, 1], [1, 1]),
    (2, 2, 8, 4, [1, 0, 0], [1, 0, 1]),
    (1, 3, 40, 0, [0, 2, 1], [1, 1, 1]),
    (1, 2, 88, 0, [0, 1], [1, 1]),        # the largest image side the quint-tap fp16 conv kernel takes (128 + 4*(D+2) <= 496)
    (1, 2, 96, 4, [0, 1, 0], [1, 0, 1]),  # the first one that falls back to the per-tap tensor-core kernel
]


@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
@pytest.mark.parametrize("B,K,D,A,sched,lw", SHAPES)
def test_edge_shapes_match_oracle(B, K, D, A, sched, lw, precision):
    T = 1 + sum(1 for s in sched if s in (1, 2))
    op3_b200.set_precision(precision)
    try:
        m, p = build_scaled(K, D, A)
        g = torch.Generator().manual_seed(5)
        frames = torch.randint(0, 256, (B, T, 3, D, D), generator=g, dtype=torch.uint8).float() / 255.0
        actions = orc.synth_actions(B, T) if A else None
        with NoiseInjector(orc.make_noise(B, K, 64, len(sched) + 1)):
            out = m.forward(frames.cuda(), actions.cuda() if A else None, None, np.array(sched, dtype=np.float64), np.array(lw))
        out[3].backward()
        grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in m.named_parameters()}
    finally:
        op3_b200.set_precision("fp32")
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    o = orc.run_schedule(q, frames, actions, sched, lw, orc.make_noise(B, K, 64, len(sched) + 1), K, D=D)
    o["total_loss"].backward()
    ograds = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in q.items()}
    tol = dict(loss=1e-5, masks=1e-5, grads=1e-4, floor=1e-6) if precision == "fp32" else dict(loss=2e-5, masks=2e-5, grads=2e-4, floor=1e-5)
    lrel = abs(out[3].item() - o["total_loss"].item()) / abs(o["total_loss"].item())
    merr = max_abs(out[1], o["masks_list"])
    print("edge B=%d K=%d D=%d %s: loss rel %.2e masks %.2e" % (B, K, D, precision, lrel, merr))
    assert lrel <= tol["loss"] and merr <= tol["masks"]
    tn = torch.sqrt(sum((v.double() ** 2).sum() for v in ograds.values())).item()
    compare_grads(grads, ograds, tn, tol["grads"], floor=tol["floor"])
