"""Two ranks over NCCL on two GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`): the data-parallel trainer step and the
sharded MPC scoring pass against their single-process results."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    import op3_b200
    from oracle import op3_oracle as orc
    from gpu_util import build_model
    from op3_b200 import parallel as par
    from op3_b200.mpc import Cost
    from op3_b200.torch.op3_modules.op3_trainer import OP3Trainer, TrainingScheduler
    op3_b200.set_precision("f16x3")
    dev = torch.device("cuda", rank)
    K, B, T = 3, 2, 3
    sched, lw = np.array([0., 0., 1., 1., 0.]), np.arange(1, 6)
    frames = orc.synth_frames(world * B, T, seed=3)
    actions = orc.synth_actions(world * B, T, seed=4)[..., [0, 1, 3, 4]].contiguous()

    # ---- data-parallel train step: every rank steps on its own B sequences; parameters stay identical and equal those of a
    # single-process step on all 2B sequences (mean of the per-rank means of equal-sized shards = the global mean)
    def one_step(fr, ac, distributed):
        m, _ = build_model(4, K, seed=2, device=dev)
        m.deterministic_sampling = True                    # no noise: the two runs see the same samples
        m.set_beta(0)
        tr = OP3Trainer(None, None, m, TrainingScheduler(4, "curriculum", T, T), fr.shape[0], lr=3e-4)
        if not distributed:                                # the trainer all-reduces whenever a process group exists
            real = dist.is_initialized
            dist.is_initialized = lambda: False
            try:
                tr.train_step(fr.to(dev), ac.to(dev), sched, lw)
            finally:
                dist.is_initialized = real
        else:
            tr.train_step(fr.to(dev), ac.to(dev), sched, lw)
        return torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    mine = one_step(frames[rank * B:(rank + 1) * B], actions[rank * B:(rank + 1) * B], True)
    both = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    assert torch.equal(both[0], both[1]), "ranks diverged after one data-parallel step"
    single = one_step(frames, actions, False)
    p0 = torch.cat([p.detach().reshape(-1) for p in build_model(4, K, seed=2, device=dev)[0].parameters()])
    step_dp, step_1 = (mine - p0).double(), (single - p0).double()
    rel = ((step_dp - step_1).norm() / step_1.norm()).item()
    assert rel <= 2e-4, "data-parallel update differs from the single-process update on the full batch: rel %.2e" % rel

    # ---- sharded MPC scoring: identical ranking on every rank, identical to the unsharded pass (bit for bit)
    m, _ = build_model(4, K, seed=7, device=dev)
    m.eval_mode = True
    m.deterministic_sampling = True
    seed_fr = orc.synth_frames(1, 2, seed=31).to(dev)
    seed_ac = orc.synth_actions(1, 2, seed=32)[:, :, [0, 1, 3, 4]].to(dev)
    info = m.batch_internal_inference(seed_fr, seed_ac, None, np.array([0., 0., 1., 0.]))
    Na = 37                                                # uneven shards: 19 + 18
    cand = orc.synth_actions(Na, 1, seed=33)[:, :, [0, 1, 3, 4]].to(dev)
    goal = {"goal_image": orc.synth_frames(1, 1, seed=34)[:, 0].to(dev)}
    cost = Cost(None, core_type="final_recon", compare_func="mse", post_process="raw", aggregate="sum")
    sc, order = par.sharded_rollout_costs(m, cost, goal, cand, info["state"], np.array([1.]))
    pred = m.batch_internal_inference(None, cand, info["state"], np.array([1.]))
    c, _ = cost.pair_costs(goal["goal_image"], pred["final_recon"].unsqueeze(1))
    sc1, order1 = c.sum(0).sort()
    assert torch.equal(order, order1) and torch.equal(sc, sc1), "sharded ranking differs from the unsharded one"
    allo = [torch.empty_like(order) for _ in range(world)]
    dist.all_gather(allo, order)
    assert torch.equal(allo[0], allo[1])
    out[rank] = (rel, int(order[0].item()))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_rank_nccl_train_step_and_sharded_mpc():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = dict(out)
    assert set(res) == {0, 1} and res[0][1] == res[1][1]
    print("two-rank NCCL: update vs single process rel %.2e, selected candidate %d" % (res[0][0], res[0][1]))
