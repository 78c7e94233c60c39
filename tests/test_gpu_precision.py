"""GPU: the tcgen05 convolution paths (3xTF32 parity mode, 1xTF32 speed mode) against the CPU oracle.

Stated tolerances.  3xTF32 (fp32-grade products; the tensor core's fp32 accumulation truncates, which leaves ~1e-5
after 100 accumulation steps per output): loss 2e-5 rel, masks 2e-5 abs, colours 5e-5 abs, parameter gradients 2e-4
rel-L2.  1xTF32 speed mode: loss 5e-4 rel, masks 2e-3 abs, gradients 5e-3 rel-L2.  Gradient
tensors whose norm is below 1e-5 (3x) / 1e-4 (1x) of the global gradient norm are compared against that floor."""
import numpy as np
import pytest
import torch

import op3_b200
from oracle import op3_oracle as orc
from gpu_util import build_model, NoiseInjector, rel_l2, max_abs, compare_grads
from test_gpu_model import CASES, run_ours

pytestmark = pytest.mark.gpu

TOL = {"tf32x3": dict(loss=2e-5, masks=2e-5, colors=5e-5, grads=2e-4, dec=3e-5, floor=1e-5),
       # fp16 hi/lo split with per-image scaling: the same 22 significand bits, same gates as 3xTF32
       "f16x3": dict(loss=2e-5, masks=2e-5, colors=5e-5, grads=2e-4, dec=3e-5, floor=1e-5),
       "tf32": dict(loss=5e-4, masks=2e-3, colors=4e-3, grads=5e-3, dec=3e-3, floor=1e-4)}


@pytest.fixture(params=["tf32x3", "tf32", "f16x3"])
def precision(request):
    op3_b200.set_precision(request.param)
    try:
        yield request.param
    finally:
        op3_b200.set_precision("fp32")


def test_decoder_forward_tcgen05(precision):
    m, p = build_model(4, 3, seed=3)
    gen = torch.Generator().manual_seed(21)
    z = torch.randn(6, 128, generator=gen) * 0.7
    logits, colors = m.decode_net(z.cuda())
    ref = orc.decoder(p, z)
    scale = ref.abs().max().item()
    e1, e2 = max_abs(colors, ref[:, :3]) / scale, max_abs(logits, ref[:, 3:4]) / scale
    print("decoder fwd %s: rel max err colors %.2e logits %.2e" % (precision, e1, e2))
    assert max(e1, e2) <= TOL[precision]["dec"]


@pytest.mark.parametrize("name", ["stack_small", "pickplace_small"])
def test_train_step_tcgen05(precision, name):
    cfg = CASES[name]
    m, p, out, grads = run_ours(name, cfg)
    B, K, T = cfg["B"], cfg["K"], cfg["T"]
    frames = orc.synth_frames(B, T)
    actions = orc.synth_actions(B, T) if cfg["action_dim"] else None
    sched = cfg["schedule"]
    lw = np.arange(1, len(sched) + 1)
    oout, ograds = orc.train_step(p, frames, actions, sched, lw, orc.make_noise(B, K, 64, len(sched) + 1), K)
    tol = TOL[precision]
    lrel = abs(out[3].item() - oout["total_loss"].item()) / abs(oout["total_loss"].item())
    merr, cerr = max_abs(out[1], oout["masks_list"]), max_abs(out[0], oout["colors_list"])
    total_norm = torch.sqrt(sum((v.double() ** 2).sum() for v in ograds.values())).item()
    report = []
    try:
        worst = compare_grads(grads, ograds, total_norm, tol["grads"], report, floor=tol["floor"])
    finally:
        print("%s %s: loss rel %.2e masks %.2e colors %.2e" % (precision, name, lrel, merr, cerr))
        print("\n".join(sorted(report, key=lambda s: -float(s.split("rel")[1].split()[0]))[:8]))
    assert lrel <= tol["loss"] and merr <= tol["masks"] and cerr <= tol["colors"]
