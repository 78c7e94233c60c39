"""GPU: stack-style MPC (mpc_stack.py:249-258): candidate AFTER-images (Na,T=1,3,D,D), no actions, no initial state,
schedule [0,0,0,0,1], K=7, ranked with Cost(subimage, mse, raw, min) (mpc_stack.py:553-558) against the goal latents
of a goal frame refined 5 times (mpc_stack.py:403-423).  Goldens come from the unmodified reference
(tests/golden/make_golden.py: mpc_stack_case, chunks of 10) at Na=40 and at BASELINE.json's Na=4096.

Gates: selected action (argmin) and its paired goal latent index-exact; sorted costs within 2e-5 relative of the cost
scale; any ranking difference only between candidates whose reference costs are tied within that tolerance."""
import os

import numpy as np
import pytest
import torch

import op3_b200
from oracle import op3_oracle as orc
from gpu_util import build_model, NoiseInjector
from helpers import load_golden, check_summary, GOLDEN
from op3_b200.mpc import Cost

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["f16x3", "fp32"])
def precision(request):
    op3_b200.set_precision(request.param)
    try:
        yield request.param
    finally:
        op3_b200.set_precision("fp32")


def _run(Na, chunk):
    K = 7
    m, p = build_model(0, K, seed=7)
    m.eval_mode = True
    m.inference_chunk = chunk
    goal_frames = orc.synth_frames(1, 1, seed=61).cuda()
    with NoiseInjector(orc.make_noise(1, K, 64, 6, seed=62)):
        goal = m.batch_internal_inference(goal_frames, None, None, np.zeros(5))
    obs = orc.synth_frames(Na, 1, seed=63).cuda()
    sched = np.array([0, 0, 0, 0, 1], dtype=np.float64)
    full = orc.make_noise(Na, K, 64, len(sched) + 1, seed=64)
    with NoiseInjector([f[i:i + chunk] for i in range(0, Na, chunk) for f in full]):
        pred = m.batch_internal_inference(obs, None, None, sched)
    cost = Cost(None, core_type="subimage", compare_func="mse", post_process="raw", aggregate="min")
    sc, order, gidx = cost.get_action_rankings(goal["state"]["post"]["samples"][0], goal["sub_images"][0], goal_frames[:, 0],
                                               pred["state"]["post"]["samples"], pred["sub_images"], pred["final_recon"],
                                               plot_actions=0)
    return goal, pred, np.asarray(sc), np.asarray(order), np.asarray(gidx)


def _check(g, goal, pred, sc, order, gidx, tol):
    check_summary(g, "goal/sub_images", goal["sub_images"], 1e-4, 1e-5)
    check_summary(g, "pred/final_recon", pred["final_recon"], 2e-4, 2e-5)
    check_summary(g, "pred/sub_images", pred["sub_images"], 2e-4, 2e-5)
    ref_sorted, ref_order, ref_gidx = g["sorted"], g["order"], g["gidx"]
    scale = float(np.abs(ref_sorted).max())
    assert np.abs(sc - ref_sorted).max() <= tol * scale, "sorted costs differ: %g" % np.abs(sc - ref_sorted).max()
    assert order[0] == ref_order[0], "selected action differs from the reference"
    assert gidx[order[0]] == ref_gidx[ref_order[0]], "paired goal latent of the selected action differs"
    # any rank difference must be between near-tied candidates
    ref_cost_of = np.empty(len(ref_order)); ref_cost_of[ref_order] = ref_sorted
    moved = np.nonzero(order != ref_order)[0]
    for r in moved:
        assert abs(ref_cost_of[order[r]] - ref_sorted[r]) <= tol * scale, \
            "rank %d: candidate %d (ref cost %g) vs reference candidate %d (cost %g)" % (
                r, order[r], ref_cost_of[order[r]], ref_order[r], ref_sorted[r])
    F = max(1, len(ref_order) // 10)                 # CEM elite set (mpc_stack.py:277: 10 %)
    boundary_gap = ref_sorted[F] - ref_sorted[F - 1] if F < len(ref_sorted) else np.inf
    if boundary_gap > 2 * tol * scale:
        assert set(order[:F].tolist()) == set(ref_order[:F].tolist()), "CEM elite set differs"
    return len(moved)


def test_stack_mpc_40_matches_reference(precision):
    g = load_golden("mpc_stack_40")
    goal, pred, sc, order, gidx = _run(40, chunk=16)
    moved = _check(g, goal, pred, sc, order, gidx, 2e-5)
    assert np.array_equal(order, g["order"]) or moved <= 2
    print("stack MPC 40 [%s]: argmin %d, %d rank differences" % (precision, order[0], moved))


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN, "mpc_stack_4096.npz")), reason="4096-candidate golden not generated")
def test_stack_mpc_4096_matches_reference(precision):
    g = load_golden("mpc_stack_4096")
    goal, pred, sc, order, gidx = _run(4096, chunk=256)
    moved = _check(g, goal, pred, sc, order, gidx, 2e-5)
    print("stack MPC 4096 [%s]: argmin %d (cost %.6g), %d rank differences between near-tied neighbours" % (
        precision, order[0], sc[0], moved))
