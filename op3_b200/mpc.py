"""MPC glue: `Cost` mirrors the ranking class of exps/stack_exps/mpc_stack.py:27-228 (== exps/pickplace_exps/
mpc_pickplace.py:36-236): core_type subimage / final_recon / latent, compare_func mse / psuedo_intersect, post_process raw /
negative_exp, aggregate sum / min; the (Na,K) x G distance reduction and the min over K run in op3_mpc_cost_ex with a fixed
reduction order, the final sort stays on the host like the reference
(it returns numpy arrays).  `shard_candidates` / `gather_costs` split CEM candidates across ranks."""
import ctypes as C

import numpy as np
import torch

from op3_b200 import _lib
from op3_b200._lib import check, ptr


class Cost:
    COMPARE = {"mse": 0, "psuedo_intersect": 1}
    POST = {"raw": 0, "negative_exp": 1}

    def __init__(self, logging_directory, core_type="subimage", compare_func="mse", post_process="raw", aggregate="sum"):
        if core_type not in ["subimage", "final_recon", "latent"]:
            raise ValueError("Invalid value of core_type: {}".format(core_type))
        if compare_func not in self.COMPARE or post_process not in self.POST or aggregate not in ("sum", "min"):
            raise KeyError((compare_func, post_process, aggregate))
        if core_type == "latent" and compare_func != "mse":
            raise KeyError(compare_func)                       # compare_latents knows 'mse' only (mpc_stack.py:73-77)
        self.logging_directory, self.core_type = logging_directory, core_type
        self.compare_func, self.post_process, self.aggregate = compare_func, post_process, aggregate

    def pair_costs(self, goal_rows, pred_rows):
        """goal (G, ...), pred (Na,K, ...) with equal trailing shapes -- sub-images (3,D,D) or latents (R,) --
        -> costs (G,Na) = min_k post(compare(goal_g, pred_a,k)), latent_idx (G,Na)  (mpc_stack.py:60-100)."""
        lib = _lib.load()
        goal = goal_rows.detach().contiguous().float()
        pred = pred_rows.detach().contiguous().float()
        Na, Kp = pred.shape[:2]
        G, n_el = goal.shape[0], goal[0].numel()
        if pred[0, 0].numel() != n_el:
            raise ValueError("goal rows %s do not match prediction rows %s" % (tuple(goal.shape[1:]), tuple(pred.shape[2:])))
        cost = torch.empty(G, Na, device=pred.device)
        argk = torch.empty(G, Na, device=pred.device, dtype=torch.int32)
        st = torch.cuda.current_stream(pred.device).cuda_stream
        check(lib.op3_mpc_cost_ex(Na, Kp, G, n_el, self.COMPARE[self.compare_func], self.POST[self.post_process], ptr(pred),
                                  ptr(goal), ptr(cost), ptr(argk), st), "op3_mpc_cost")
        return cost, argk.long()

    def get_action_rankings(self, goal_latents, goal_latents_recon, goal_image, pred_latents, pred_latents_recon,
                            pred_image, image_suffix="", plot_actions=8):
        if self.core_type == "final_recon":                   # pretend K = 1 (mpc_stack.py:46-48)
            goal_latents_recon = goal_image
            pred_latents_recon = pred_image.unsqueeze(1)
        if self.core_type == "latent":                        # one cost row per goal latent RECON (mpc_stack.py:116-122)
            costs, latent_idxs = self.pair_costs(goal_latents[:goal_latents_recon.shape[0]], pred_latents)
        else:
            costs, latent_idxs = self.pair_costs(goal_latents_recon, pred_latents_recon)
        if self.aggregate == "sum":
            sorted_costs, order = costs.sum(0).sort()
            return sorted_costs.cpu().numpy(), order.cpu().numpy(), np.zeros(len(sorted_costs))
        min_costs, goal_idx = costs.min(0)
        sorted_costs, order = min_costs.sort()
        return sorted_costs.cpu().numpy(), order.cpu().numpy(), goal_idx.cpu().numpy()


def cem_refit(actions, order, n_elite, min_std=0.0):
    """Mean and (population) std of the `n_elite` best candidates, on the device (the reference does this in numpy on
    the host after a D2H copy: mpc_stack.py:283-288 with min_std 0, mpc_pickplace.py:376-381 with min_std 0.05).
    actions (Na, ...) float32 cuda, order (>= n_elite,) int64 candidate indices sorted by cost."""
    lib = _lib.load()
    a = actions.detach().contiguous().float()
    if a.device.type != "cuda":
        raise RuntimeError("op3_b200: cem_refit needs CUDA tensors (no CPU fallback)")
    Na, TA = a.shape[0], a[0].numel()
    order = order.to(device=a.device, dtype=torch.int64).contiguous()
    if not 0 < n_elite <= min(Na, order.numel()):
        raise ValueError("n_elite=%d out of range for %d candidates" % (n_elite, Na))
    mean = torch.empty(a.shape[1:], device=a.device)
    std = torch.empty(a.shape[1:], device=a.device)
    st = torch.cuda.current_stream(a.device).cuda_stream
    check(lib.op3_cem_refit(Na, TA, int(n_elite), ptr(a), ptr(order), float(min_std), ptr(mean), ptr(std), st), "op3_cem_refit")
    return mean, std


def cem_resample(mean, std, noise, low=None, high=None):
    """actions = clip(mean + std * noise, low, high): env.sample_multiple_action_gaussian
    (op3/envs/blocks/mujoco/block_pick_and_place.py:390-407) with the standard-normal `noise` (Na, *mean.shape) supplied
    by the caller (same noise on every rank => identical candidates without a broadcast)."""
    lib = _lib.load()
    noise = noise.detach().contiguous().float()
    if noise.device.type != "cuda":
        raise RuntimeError("op3_b200: cem_resample needs CUDA tensors (no CPU fallback)")
    Na, TA = noise.shape[0], mean.numel()
    if tuple(noise.shape[1:]) != tuple(mean.shape) or tuple(std.shape) != tuple(mean.shape):
        raise ValueError("noise %s does not match mean %s / std %s" % (tuple(noise.shape), tuple(mean.shape), tuple(std.shape)))
    dev = noise.device
    as_vec = lambda t: None if t is None else torch.as_tensor(t, dtype=torch.float32, device=dev).expand(mean.shape).contiguous()
    lo, hi = as_vec(low), as_vec(high)
    out = torch.empty_like(noise)
    st = torch.cuda.current_stream(dev).cuda_stream
    check(lib.op3_cem_resample(Na, TA, ptr(mean.contiguous().float()), ptr(std.contiguous().float()), ptr(noise), ptr(lo),
                               ptr(hi), ptr(out), st), "op3_cem_resample")
    return out


def cem_plan(model, cost, goal_info, initial_hidden_state, actions, cem_steps, noise_fn, low=None, high=None,
             elite_frac=0.1, min_std=0.05, schedule=None):
    """Device-resident CEM loop of Stage3_CEM._cem (mpc_pickplace.py:362-374, action_type None): roll the candidate
    action sequences out through dynamics + decoder, rank them with `cost`, refit a Gaussian on the best 10 % and
    resample -- candidates, costs, ranking and refit stay on the GPU; with torch.distributed initialised the candidate
    set is sharded across ranks (op3_b200.parallel.sharded_rollout_costs) and every rank refits the same distribution.
    actions (Na,T,A) initial candidates (identical on every rank); noise_fn(step) -> (Na,T,A) standard-normal noise
    (identical on every rank).  Returns (best action sequence (T,A), its cost, final candidates, final order)."""
    from op3_b200 import parallel as par
    Na, T = actions.shape[0], actions.shape[1]
    schedule = [1] * T if schedule is None else schedule
    n_elite = int(Na * elite_frac)
    best, best_cost, order = None, None, None
    for i in range(cem_steps):
        sorted_costs, order = par.sharded_rollout_costs(model, cost, goal_info, actions, initial_hidden_state, schedule)
        best, best_cost = actions[order[0]].clone(), sorted_costs[0]
        if i + 1 < cem_steps:
            mean, std = cem_refit(actions, order, n_elite, min_std)
            actions = cem_resample(mean, std, noise_fn(i), low, high)
    return best, best_cost, actions, order
