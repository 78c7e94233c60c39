"""MPC rollout throughput (BASELINE.json config 5, pickplace form): 4096 candidate actions rolled out of one B=1 state
through dynamics + decoder (schedule [1]), costs against a goal image, argsort.  Reported, not part of bench.py's metric.
usage: python tools/bench_mpc.py [precision]"""
import json
import sys
import os
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import op3_b200
from op3_b200.mpc import Cost
from gpu_util import build_model
from oracle import op3_oracle as orc   # only for the synthetic input generators

prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
op3_b200.set_precision(prec)
K, Na = 4, 4096
m, p = build_model(4, K, seed=7)
m.eval_mode = True
frames = orc.synth_frames(1, 2, seed=31).cuda()
acts = orc.synth_actions(1, 2, seed=32)[:, :, [0, 1, 3, 4]].cuda()
info = m.batch_internal_inference(frames, acts, None, np.array([0., 0., 1., 0.]))
cand = orc.synth_actions(Na, 1, seed=33)[:, :, [0, 1, 3, 4]].cuda()
goal = orc.synth_frames(1, 1, seed=34)[:, 0].cuda()
cost = Cost(None, core_type="final_recon", aggregate="sum")


def rollout():
    pred = m.batch_internal_inference(None, cand, info["state"], np.array([1.]))
    c, _ = cost.pair_costs(goal, pred["final_recon"].unsqueeze(1))
    return torch.argsort(c[0], stable=True)


for _ in range(3):
    rollout()
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
ev0.record()
for _ in range(n):
    order = rollout()
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / n
print(json.dumps({"metric": "mpc_candidates_per_sec", "value": Na / (ms * 1e-3), "ms_per_rollout": ms, "candidates": Na, "K": K,
                  "precision": prec, "best": int(order[0])}))
