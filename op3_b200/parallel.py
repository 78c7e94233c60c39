"""One-process-per-GPU data parallelism for the OP3 path (replaces nn.DataParallel, exps/train_op3/train_op3.py:99-100).

Training: samples are independent (no BatchNorm; every LayerNorm is per sample), so the batch is split across ranks
and the only exchange is ONE all-reduce of the flat gradient buffer per step (3.5-3.8 MB, latency bound), followed by
the identical fused clip+Adam on every rank.  MPC / CEM: candidates are independent; each rank rolls out and scores its
shard and the Na fp32 costs are all-gathered so that every rank ranks the same list.
Backend: NCCL over NVLink on GPUs; the same code runs on gloo for the CPU tests of this host logic."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank=None, world_size=None):
    """Contiguous [start, end) of `n` items owned by `rank`; the first n % world ranks get one extra item."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_mean_(flat_grad):
    """In-place mean over ranks of a flat gradient buffer (== DataParallel's mean of equal-size replicas)."""
    _, w = world()
    if w > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        flat_grad.div_(w)
    return flat_grad


def all_gather_rows(local, n_total):
    """Concatenate per-rank row blocks (split with shard_range) into the full (n_total, ...) tensor on every rank."""
    r, w = world()
    if w == 1:
        return local
    base, extra = divmod(n_total, w)
    width = base + (1 if extra else 0)
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(w)]
    dist.all_gather(parts, padded)
    out = []
    for k, p in enumerate(parts):
        s, e = shard_range(n_total, k, w)
        out.append(p[:e - s])
    return torch.cat(out)


def sharded_rollout_costs(model, cost, goal_info, actions, initial_hidden_state, schedule, obs=None):
    """CEM scoring with the candidate set split across ranks: pickplace style (mpc_pickplace.py:321-340: candidate
    actions (Na,T,A) rolled out of a B=1 state) or stack style (mpc_stack.py:249-258: candidate after-images `obs`
    (Na,T,3,D,D), no actions, no initial state).  The candidates must be identical on every rank (same seed or
    broadcast).  Returns (sorted_costs, order) computed from the gathered per-candidate costs, identical on every rank."""
    Na = actions.shape[0] if actions is not None else obs.shape[0]
    s, e = shard_range(Na)
    pred = model.batch_internal_inference(obs=None if obs is None else obs[s:e], actions=None if actions is None else actions[s:e],
                                          initial_hidden_state=initial_hidden_state, schedule=schedule)
    if cost.core_type == "final_recon":
        goal, p = goal_info["goal_image"], pred["final_recon"].unsqueeze(1)
    else:
        goal, p = goal_info["sub_images"][0], pred["sub_images"]
    c, _ = cost.pair_costs(goal, p)                       # (G, Na_local)
    per_candidate = c.sum(0) if cost.aggregate == "sum" else c.min(0)[0]
    full = all_gather_rows(per_candidate, Na)
    sorted_costs, order = full.sort()
    return sorted_costs, order
