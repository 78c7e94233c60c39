"""MPC glue: `Cost` mirrors the ranking class of exps/stack_exps/mpc_stack.py:27-228 (== exps/pickplace_exps/
mpc_pickplace.py:36-236) for compare_func='mse', post_process='raw'; the (Na,K) x G squared-error reduction and the
min over K run in op3_mpc_cost with a fixed reduction order, the final sort stays on the host like the reference
(it returns numpy arrays).  `shard_candidates` / `gather_costs` split CEM candidates across ranks."""
import ctypes as C

import numpy as np
import torch

from op3_b200 import _lib
from op3_b200._lib import check, ptr


class Cost:
    def __init__(self, logging_directory, core_type="subimage", compare_func="mse", post_process="raw", aggregate="sum"):
        if core_type not in ["subimage", "final_recon", "latent"]:
            raise ValueError("Invalid value of core_type: {}".format(core_type))
        if core_type == "latent" or compare_func != "mse" or post_process != "raw":
            raise NotImplementedError("only core_type in {subimage, final_recon} with mse/raw is built")
        if aggregate not in ("sum", "min"):
            raise KeyError(aggregate)
        self.logging_directory, self.core_type = logging_directory, core_type
        self.compare_func, self.post_process, self.aggregate = compare_func, post_process, aggregate

    def pair_costs(self, goal_latents_recon, pred_latents_recon):
        """(G,3,D,D), (Na,K,3,D,D) -> costs (G,Na) = min_k mse, latent_idx (G,Na)."""
        lib = _lib.load()
        goal = goal_latents_recon.detach().contiguous().float()
        pred = pred_latents_recon.detach().contiguous().float()
        Na, Kp, _, D, _ = pred.shape
        G = goal.shape[0]
        cost = torch.empty(G, Na, device=pred.device)
        argk = torch.empty(G, Na, device=pred.device, dtype=torch.int32)
        st = torch.cuda.current_stream(pred.device).cuda_stream
        check(lib.op3_mpc_cost(Na, Kp, G, D, ptr(pred), ptr(goal), ptr(cost), ptr(argk), st), "op3_mpc_cost")
        return cost, argk.long()

    def get_action_rankings(self, goal_latents, goal_latents_recon, goal_image, pred_latents, pred_latents_recon,
                            pred_image, image_suffix="", plot_actions=8):
        if self.core_type == "final_recon":                   # pretend K = 1 (mpc_stack.py:46-48)
            goal_latents_recon = goal_image
            pred_latents_recon = pred_image.unsqueeze(1)
        costs, latent_idxs = self.pair_costs(goal_latents_recon, pred_latents_recon)
        if self.aggregate == "sum":
            sorted_costs, order = costs.sum(0).sort()
            return sorted_costs.cpu().numpy(), order.cpu().numpy(), np.zeros(len(sorted_costs))
        min_costs, goal_idx = costs.min(0)
        sorted_costs, order = min_costs.sort()
        return sorted_costs.cpu().numpy(), order.cpu().numpy(), goal_idx.cpu().numpy()
