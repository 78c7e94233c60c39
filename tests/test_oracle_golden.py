"""CPU: pin the oracle restatement against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Tolerances are the fp32 noise floor measured in SURVEY.md section 8c
(reference fp32 vs its own fp64: loss 5e-8 rel, masks 5e-7 abs, grads 5e-6..2e-5 rel-L2) with head-room."""
import numpy as np
import pytest
import torch

from oracle import op3_oracle as orc
from helpers import load_golden, check_summary, check_ranking, COST_OPTION_CASES

CASES = [("stack_small", 0), ("pickplace_small", 4), ("stack_c1", 0), ("pickplace_nextstep", 4)]


@pytest.mark.parametrize("name,action_dim", CASES)
def test_train_step_matches_reference(name, action_dim):
    g = load_golden(name)
    B, K, T = int(g["B"]), int(g["K"]), int(g["T"])
    sched, lw = g["schedule"], g["loss_schedule"]
    p = orc.make_params(action_dim, seed=0)
    frames = orc.synth_frames(B, T)
    actions = orc.synth_actions(B, T) if action_dim else None
    noise = orc.make_noise(B, K, 64, len(sched) + 1)
    out, grads = orc.train_step(p, frames, actions, sched, lw, noise, K)
    for key, ours in [("total_loss", out["total_loss"]), ("kle_loss", out["kle_loss"]),
                      ("clog_prob", out["clog_prob"]), ("mse", out["mse"])]:
        assert abs(ours.item() - float(g[key])) <= 2e-6 * abs(float(g[key])), key
    if "masks_list" in g.files:
        assert np.abs(out["masks_list"].detach().numpy() - g["masks_list"]).max() <= 5e-6
        assert np.abs(out["final_recon"].detach().numpy() - g["final_recon"]).max() <= 5e-6
    check_summary(g, "masks_list", out["masks_list"], 1e-5, 5e-6)
    check_summary(g, "colors_list", out["colors_list"], 1e-5, 5e-6)
    st = out["state"]["post"]
    for k in ["lambdas1", "lambdas2", "samples", "deter_state"]:
        check_summary(g, "state/" + k, st[k], 2e-5, 1e-6)
    check_summary(g, "state/h", st["extra_info"][0], 2e-5, 1e-6)
    check_summary(g, "state/c", st["extra_info"][1], 2e-5, 1e-6)
    for k, v in grads.items():
        check_summary(g, "grad/" + k, v, 2e-4, 1e-7 * float(g["grad_total_norm"]), what=name)
    # clip + Adam (op3_trainer.py:188-189)
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v2 = {k: torch.zeros_like(v) for k, v in p.items()}
    q = {k: v.clone() for k, v in p.items()}
    total = orc.clip_adam_step(q, grads, m, v2, step=1)
    assert abs(total - float(g["grad_total_norm"])) <= 2e-5 * total
    for k in p:
        # first Adam step is +-lr*sign(g) for all but vanishing gradients -> compare absolutely in units of lr
        d = (q[k] - p[k]).double().reshape(-1)
        idx = torch.from_numpy(g["adam/" + k + "/idx"])
        assert (d[idx] - torch.from_numpy(g["adam/" + k + "/val"])).abs().max().item() <= 3e-4 * 0.05, k


def test_modules_match_reference():
    g = load_golden("modules")
    p = orc.make_params(4, seed=3)
    out = orc.decoder(p, torch.from_numpy(g["dec/z"]))
    check_summary(g, "dec/colors", out[:, :3], 1e-5, 1e-6)
    check_summary(g, "dec/logits", out[:, 3:4], 1e-5, 1e-6)
    gen = torch.Generator().manual_seed(21)
    _ = torch.randn(6, 128, generator=gen)
    a = torch.randn(6, 15, 64, 64, generator=gen)
    h = torch.randn(6, 128, generator=gen) * 0.3
    c = torch.randn(6, 128, generator=gen) * 0.3
    extra = torch.randn(6, 320, generator=gen)
    l1, l2, h2, c2 = orc.refinement_net(p, a, h, c, extra)
    for k, v in [("l1", l1), ("l2", l2), ("h2", h2), ("c2", c2)]:
        assert np.abs(v.numpy() - g["ref/" + k]).max() <= 2e-6, k
    z = torch.from_numpy(g["dyn/z"])
    acts = torch.from_numpy(g["dyn/actions"])
    d, l1, l2, sa, ia, dl = orc.dynamics_net(p, z, acts, 3, return_probes=True)
    for k, v in [("deter", d), ("l1", l1), ("l2", l2), ("act_att", sa), ("pair_att", ia), ("deltas", dl)]:
        assert np.abs(v.numpy() - g["dyn/" + k]).max() <= 5e-6 * max(1.0, np.abs(g["dyn/" + k]).max()), k


def test_mpc_rollout_and_ranking_match_reference():
    g = load_golden("mpc")
    K, Na = int(g["K"]), int(g["Na"])
    p = {k: v.requires_grad_(True) for k, v in orc.make_params(4, seed=7).items()}
    frames = orc.synth_frames(1, 2, seed=31)
    acts = orc.synth_actions(1, 2, seed=32)[:, :, [0, 1, 3, 4]]
    sched = [0, 0, 1, 0]
    out = orc.run_schedule(p, frames, acts, sched, [0] * 4, orc.make_noise(1, K, 64, 5, seed=40), K)
    st = out["state"]
    check_summary(g, "state/samples", st["post"]["samples"], 2e-5, 1e-6)
    goal_sub = (out["colors_list"][:, -1] * out["masks_list"][:, -1])[0].detach()
    # replicate_state (op3_model.py:475-494): note prior.lambdas2 := prior.lambdas1 (trap T7)
    rep = lambda t: t.detach().repeat(Na, 1, 1)
    state = {"prior": {"lambdas1": rep(st["prior"]["lambdas1"]), "lambdas2": rep(st["prior"]["lambdas1"])},
             "post": {"deter_state": rep(st["post"]["deter_state"]), "lambdas1": rep(st["post"]["lambdas1"]),
                      "lambdas2": rep(st["post"]["lambdas2"]), "samples": rep(st["post"]["samples"]),
                      "extra_info": [st["post"]["extra_info"][0].detach().repeat(Na, 1),
                                     st["post"]["extra_info"][1].detach().repeat(Na, 1)]}}
    cand = orc.synth_actions(Na, 1, seed=33)[:, :, [0, 1, 3, 4]]
    with torch.no_grad():
        pred = orc.run_schedule(p, None, cand, [1], [0], orc.make_noise(Na, K, 64, 1, seed=50), K, init_state=state)
    sub = pred["colors_list"][:, -1] * pred["masks_list"][:, -1]
    check_summary(g, "pred/final_recon", pred["final_recon"], 2e-5, 1e-6)
    check_summary(g, "pred/sub_images", sub, 2e-5, 1e-6)
    goal_image = orc.synth_frames(1, 1, seed=34)[:, 0]
    for core, agg in [("final_recon", "sum"), ("subimage", "min"), ("subimage", "sum")]:
        sc, order, gidx = orc.mpc_cost(goal_sub, goal_image, sub, pred["final_recon"], core, agg)
        key = "cost/%s_%s/" % (core, agg)
        assert np.array_equal(order.numpy(), g[key + "order"]), key          # index-exact ranking
        assert np.abs(sc.numpy() - g[key + "sorted"]).max() <= 1e-6
        if agg == "min":
            assert np.array_equal(gidx.numpy(), g[key + "gidx"])
    # the remaining Cost options (mpc_stack.py:60-100): latent distances, 'psuedo_intersect', 'negative_exp'
    for core, cmp_, post, agg in COST_OPTION_CASES:
        sc, order, gidx = orc.mpc_cost(goal_sub, goal_image, sub, pred["final_recon"], core, agg, cmp_, post,
                                       goal_latents=st["post"]["samples"][0].detach(), pred_latents=pred["state"]["post"]["samples"])
        key = "costx/%s_%s_%s_%s/" % (core, cmp_, post, agg)
        check_ranking(sc.numpy(), order.numpy(), g[key + "sorted"], g[key + "order"], exact=cmp_ == "mse", what=key)
