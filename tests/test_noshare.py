""" the *_No_Sharing registry entries -- DecoderNetwork_v2_No_Sharing (decoder_network.py:92),
RefinementNetwork_v2_No_Sharing (refinement_network.py:235), PhysicsNetwork_v2_No_Sharing (physics_network.py:192) -- against
fixtures of the unmodified reference (tests/golden/make_golden.py noshare): the CPU oracle on CPU, the CUDA engine on the GPU.
The reference concatenates the K per-slot outputs slot-major and reads them (b, k)-major again; both sides keep that."""
import numpy as np
import pytest
import torch

from oracle import op3_oracle as orc
from helpers import load_golden, check_summary

CASES = ["noshare_all", "noshare_dec_ref", "noshare_dyn"]
TYPES = {"decode_net": ("decoder_model_type", "reg_no_share"),
         "refinement_net": ("refinement_model_type", "size_dependent_conv_no_share"),
         "dynamics_net": ("dynamics_model_type", "reg_ac32_no_share")}


def _inputs(g):
    B, K, T, A = int(g["B"]), int(g["K"]), int(g["T"]), int(g["action_dim"])
    nets = tuple(str(n) for n in g["nets"])
    sched, lw = g["schedule"], g["loss_schedule"]
    frames = orc.synth_frames(B, T)
    actions = orc.synth_actions(B, T) if A else None
    return B, K, T, A, nets, sched, lw, frames, actions, orc.make_noise(B, K, 64, len(sched) + 1)


@pytest.mark.parametrize("name", CASES)
def test_oracle_no_sharing_matches_reference(name):
    g = load_golden(name)
    B, K, T, A, nets, sched, lw, frames, actions, noise = _inputs(g)
    p = orc.make_params_no_share(A, K, nets=nets, seed=0)
    out, grads = orc.train_step(p, frames, actions, sched, lw, noise, K)
    for key in ("total_loss", "kle_loss", "clog_prob", "mse"):
        assert abs(out[key].item() - float(g[key])) <= 2e-6 * abs(float(g[key])), key
    assert np.abs(out["masks_list"].detach().numpy() - g["masks_list"]).max() <= 5e-6
    assert np.abs(out["final_recon"].detach().numpy() - g["final_recon"]).max() <= 5e-6
    st = out["state"]["post"]
    for k in ["lambdas1", "lambdas2", "samples", "deter_state"]:
        check_summary(g, "state/" + k, st[k], 2e-5, 1e-6)
    check_summary(g, "state/h", st["extra_info"][0], 2e-5, 1e-6)
    assert set(grads) == {k[len("grad/"):-len("/norm")] for k in g.files if k.startswith("grad/") and k.endswith("/norm")}
    for k, v in grads.items():
        check_summary(g, "grad/" + k, v, 2e-4, 1e-7 * float(g["grad_total_norm"]), what=name)


def test_unrunnable_registry_entries_fail_like_the_reference():
    """refinement_network.py:15-62: 'large_reg' and 'large_sequence' raise RuntimeError in RefinementNetwork.__init__ of the
    reference (AvgPool2d(64) on a 52x52 / 1x1 map); the mirror keeps the registry keys and fails the same way."""
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    g = load_golden("registry_unrunnable")
    for key in ("large_reg", "large_sequence"):
        assert str(g[key]).startswith("RuntimeError: Given input size")
        v = dict(refinement_model_type=key, decoder_model_type="reg", dynamics_model_type="reg_ac32", sto_repsize=64,
                 det_repsize=64, K=2, extra_args=dict(beta=1e-2, deterministic_sampling=False))
        with pytest.raises(RuntimeError, match="Output size is too small"):
            create_model_v2(v, 64, 64, 4)


def test_no_sharing_state_dict_keys_match_reference_layout():
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    K = 3
    v = dict(refinement_model_type="size_dependent_conv_no_share", decoder_model_type="reg_no_share",
             dynamics_model_type="reg_ac32_no_share", sto_repsize=64, det_repsize=64, K=K,
             extra_args=dict(beta=1e-2, deterministic_sampling=False))
    m = create_model_v2(v, 64, 64, 4)
    p = orc.make_params_no_share(4, K, seed=0)
    assert set(m.state_dict().keys()) == set(p.keys())
    m.load_state_dict(p, strict=True)
    for k, t in m.state_dict().items():
        assert tuple(t.shape) == tuple(p[k].shape), k


def _build(nets, K, A):
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    v = dict(refinement_model_type="size_dependent_conv", decoder_model_type="reg", dynamics_model_type="reg_ac32",
             sto_repsize=64, det_repsize=64, K=K, extra_args=dict(beta=1e-2, deterministic_sampling=False))
    for n in nets:
        v[TYPES[n][0]] = TYPES[n][1]
    m = create_model_v2(v, 64, 64, A)
    m.load_state_dict(orc.make_params_no_share(A, K, nets=nets, seed=0), strict=True)
    return m.cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "f16x3"])
@pytest.mark.parametrize("name", CASES)
def test_engine_no_sharing_matches_reference(name, precision):
    import op3_b200
    from gpu_util import NoiseInjector
    g = load_golden(name)
    B, K, T, A, nets, sched, lw, frames, actions, noise = _inputs(g)
    op3_b200.set_precision(precision)
    try:
        m = _build(nets, K, A)
        with NoiseInjector(noise):
            out = m.forward(frames.cuda(), None if actions is None else actions.cuda(), None, sched, lw)
        out[3].backward()
        torch.cuda.synchronize()
    finally:
        op3_b200.set_precision("fp32")
    colors, masks, recon, total, kle, clog, mse, state = out
    lt, mt, gt = (1e-5, 1e-5, 2e-4) if precision == "fp32" else (2e-5, 2e-5, 2e-4)
    for key, ours in [("total_loss", total), ("kle_loss", kle), ("clog_prob", clog), ("mse", mse)]:
        ref = float(g[key])
        assert abs(ours.item() - ref) <= lt * abs(ref) + 1e-7, "%s: %r vs golden %r" % (key, ours.item(), ref)
    assert np.abs(masks.cpu().numpy() - g["masks_list"]).max() <= mt
    assert np.abs(recon.cpu().numpy() - g["final_recon"]).max() <= 2 * mt
    for k in ["lambdas1", "lambdas2", "samples", "deter_state"]:
        check_summary(g, "state/" + k, state["post"][k], 1e-4, 1e-5)
    check_summary(g, "state/h", state["post"]["extra_info"][0], 1e-4, 1e-5)
    gn = float(g["grad_total_norm"])
    for k, v in m.named_parameters():
        assert v.grad is not None, "no gradient for %s" % k
        check_summary(g, "grad/" + k, v.grad, gt, 2e-6 * gn, what=name)


@pytest.mark.gpu
def test_no_sharing_module_forwards_match_oracle():
    """The standalone nn.Module entry points of the three *_No_Sharing classes (slot-major outputs)."""
    K, A, B = 3, 4, 2
    m = _build(orc.NETS, K, A)
    p = orc.make_params_no_share(A, K, seed=0)
    gen = torch.Generator().manual_seed(5)
    z = torch.randn(B * K, 128, generator=gen) * 0.7
    logits, colors = m.decode_net(z.cuda())
    ref = orc.decoder(p, z)
    assert (colors.cpu() - ref[:, :3]).abs().max().item() <= 1e-5 and (logits.cpu() - ref[:, 3:4]).abs().max().item() <= 1e-5
    acts = orc.synth_actions(B, 1, seed=4)[:, 0][:, [0, 1, 3, 4]].unsqueeze(1).repeat(1, K, 1).reshape(B * K, 4)
    d, l1, l2 = m.dynamics_net(z.cuda(), acts.cuda())
    rd, r1, r2 = orc.dynamics_net(p, z, acts, K)
    for a, b in ((d, rd), (l1, r1), (l2, r2)):
        assert (a.cpu() - b).abs().max().item() <= 1e-5 * max(1.0, b.abs().max().item())
    a_in = torch.randn(B * K, 15, 64, 64, generator=gen)
    h, c = torch.randn(B * K, 128, generator=gen) * 0.3, torch.randn(B * K, 128, generator=gen) * 0.3
    extra = torch.randn(B * K, 320, generator=gen)
    o = m.refinement_net(a_in.cuda(), h.cuda(), c.cuda(), extra.cuda())
    r = orc.refinement_net(p, a_in, h, c, extra)
    assert tuple(o[2].shape) == (K, B, 128)                         # (K, bs, lstm): refinement_network.py:304-305
    for a, b in zip(o, r):
        assert (a.cpu().reshape(b.shape) - b).abs().max().item() <= 1e-5
