"""MPC rollout throughput (BASELINE.json config 5, pickplace form): 4096 candidate actions rolled out of one B=1 state
through dynamics + decoder (schedule [1]), costs against a goal image, argsort.  Reported, not part of bench.py's metric.
usage: python tools/bench_mpc.py [precision]"""
import json
import sys
import os
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import op3_b200
from op3_b200.mpc import Cost
from op3_b200.torch.op3_modules.op3_model import create_model_v2

prec = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
op3_b200.set_precision(prec)
K, Na = 4, 4096
torch.manual_seed(7)
variant = dict(refinement_model_type="size_dependent_conv", decoder_model_type="reg", dynamics_model_type="reg_ac32",
               sto_repsize=64, det_repsize=64, K=K, extra_args=dict(beta=1e-2, deterministic_sampling=False))
m = create_model_v2(variant, 64, 64, 4).cuda()
m.eval_mode = True
g = torch.Generator().manual_seed(31)
frames = (torch.randint(0, 256, (1, 2, 3, 64, 64), generator=g, dtype=torch.uint8).float() / 255.0).cuda()
acts = (torch.rand(1, 2, 4, generator=g) * 2 - 1).cuda()
info = m.batch_internal_inference(frames, acts, None, np.array([0., 0., 1., 0.]))
cand = (torch.rand(Na, 1, 4, generator=g) * 2 - 1).cuda()
goal = (torch.randint(0, 256, (1, 3, 64, 64), generator=g, dtype=torch.uint8).float() / 255.0).cuda()
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
