"""GPU: the reference-facing surface beyond forward/backward -- attention probes at model level
(op3_model.py:645-698), Mlp.forward (networks.py:63-74), checkpoint round trips (op3_trainer.py:145-146,
mpc_stack.py:465-473) and optimizer-state resume."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import op3_oracle as orc
from helpers import load_golden
from gpu_util import build_model, NoiseInjector, max_abs

pytestmark = pytest.mark.gpu


def test_get_all_activation_values_match_reference_golden():
    """Golden from the unmodified reference (make_golden.activation_case): 23 probe actions on a B=1 state."""
    g = load_golden("activations")
    K, Np = int(g["K"]), int(g["Np"])
    m, p = build_model(4, K, seed=7)
    m.eval_mode = True
    frames = orc.synth_frames(1, 2, seed=31).cuda()
    acts = orc.synth_actions(1, 2, seed=32)[:, :, [0, 1, 3, 4]].cuda()
    with NoiseInjector(orc.make_noise(1, K, 64, 5, seed=40)):
        info = m.batch_internal_inference(frames, acts, None, np.array([0., 0., 1., 0.]))
    probe = orc.synth_actions(Np, 1, seed=35)[:, 0][:, [0, 1, 3, 4]].cuda()
    sa, ia, dv, ld = m.get_all_activation_values(info["state"], probe)
    assert sa.shape == (Np, K) and ia.shape == (Np, K, K - 1) and dv.shape == (Np, K) and ld.shape == (Np, K)
    assert max_abs(sa, torch.from_numpy(g["state_action"])) <= 2e-5
    assert max_abs(ia, torch.from_numpy(g["interaction"])) <= 2e-5
    for ours, key in [(dv, "deltas"), (ld, "lambda_deltas")]:
        ref = torch.from_numpy(g[key])
        assert max_abs(ours, ref) <= 1e-4 * ref.abs().max().item() + 1e-5, key
    sa1 = m.get_activation_values(info["state"], probe)
    assert max_abs(sa1, torch.from_numpy(g["state_action_only"])) <= 2e-5


@pytest.mark.parametrize("hidden,out_act", [(F.relu, None), (torch.nn.ELU(), torch.nn.Sigmoid()), (torch.tanh, None)])
def test_mlp_forward_matches_torch(hidden, out_act):
    from op3_b200.torch.networks import Mlp, identity
    torch.manual_seed(3)
    mlp = Mlp((96, 40), 7, 33, hidden_activation=hidden, output_activation=out_act or identity)
    x = torch.randn(5, 11, 33)
    h = x
    for fc in mlp.fcs:
        h = hidden(F.linear(h, fc.weight, fc.bias))
    pre = F.linear(h, mlp.last_fc.weight, mlp.last_fc.bias)
    want = out_act(pre) if out_act else pre
    mlp = mlp.cuda()
    got, got_pre = mlp(x.cuda(), return_preactivations=True)
    assert got.shape == (5, 11, 7)
    assert max_abs(got, want.detach()) <= 2e-6 and max_abs(got_pre, pre.detach()) <= 2e-6
    assert max_abs(mlp(x.cuda()), want.detach()) <= 2e-6
    with pytest.raises(RuntimeError, match="CUDA-only"):
        mlp(x)


def test_save_model_reload_and_optimizer_resume(tmp_path):
    """save_model writes `<prefix>_params.pkl` with `module.` keys (the reference's DataParallel checkpoints,
    op3_trainer.py:145-146); reloading it the way mpc_stack.py:465-473 does gives the same model; restoring the
    optimizer state_dict continues the Adam trajectory bit for bit."""
    from op3_b200.torch.op3_modules.op3_trainer import OP3Trainer, TrainingScheduler
    B, K, T, sched = 2, 3, 2, np.array([0., 0., 1.])
    lw = np.arange(1, 4)
    frames = orc.synth_frames(B, T).cuda()

    def make():
        m, _ = build_model(0, K, seed=0)
        return OP3Trainer(None, None, m, TrainingScheduler(2, "single_step_physics", None, T), B, lr=3e-4)

    def step(tr, i):
        with NoiseInjector(orc.make_noise(B, K, 64, 4, seed=100 + i)):
            return [float(v) for v in tr.train_step(frames, None, sched, lw)]

    a = make()
    for i in range(2):
        step(a, i)
    a.save_model("ckpt", directory=str(tmp_path))
    opt_state = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in a.optimizer.state_dict().items()}
    want = [step(a, i) for i in (2, 3)]

    # reload like the MPC scripts: strip `module.`, load_state_dict strict
    sd = torch.load(os.path.join(str(tmp_path), "ckpt_params.pkl"), map_location="cpu")
    assert all(k.startswith("module.") for k in sd)
    b = make()
    b.model.load_state_dict({k[len("module."):]: v for k, v in sd.items()}, strict=True)
    b.optimizer.load_state_dict(opt_state)
    got = [step(b, i) for i in (2, 3)]
    assert got == want, "resumed trajectory differs: %r vs %r" % (got, want)
    for (k, pa), (_, pb) in zip(a.model.named_parameters(), b.model.named_parameters()):
        assert torch.equal(pa, pb), k


def test_train_vae_loop_body_with_dataparallel_wrapper(tmp_path):
    """The body of train_vae (exps/train_op3/train_op3.py:96-116) on synthetic data: float TensorDataset + BlocksDataset
    (as load_dataset builds them), create_model_v2, nn.DataParallel on ONE visible device, OP3Trainer.train_epoch /
    test_epoch / save_model."""
    from torch.utils.data.dataset import TensorDataset
    from op3_b200.torch.data_management.dataset import BlocksDataset
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    from op3_b200.torch.op3_modules.op3_trainer import OP3Trainer, TrainingScheduler
    from gpu_util import VARIANT
    torch.manual_seed(0)
    feats = torch.randint(0, 256, (12, 5, 3, 64, 64)).float()                 # (s,T,3,D,D), values 0..255
    actions = torch.randn(12, 5, 6)                                             # pickplace: 6-dim, cols [0,1,3,4] used
    train = BlocksDataset(TensorDataset(feats, actions), batchsize=4)
    train.action_dim = 4                                                        # train_op3.py:59-60
    test = BlocksDataset(TensorDataset(feats[:4], actions[:4]), batchsize=4)
    m = create_model_v2(dict(VARIANT, K=3), 64, 64, action_dim=train.action_dim)
    m = torch.nn.DataParallel(m, device_ids=[0])
    m.cuda()
    t = OP3Trainer(train, test, m, TrainingScheduler(seed_steps=2, schedule_type="curriculum", max_T=5, T=3), batch_size=4, lr=3e-4)
    before = {k: v.detach().clone() for k, v in m.module.named_parameters()}
    stats = t.train_epoch(0)
    assert np.isfinite(stats["train/loss"]) and np.isfinite(stats["train/mse"]) and stats["train/epoch"] == 0
    assert any(not torch.equal(v, before[k]) for k, v in m.module.named_parameters())
    ts = t.test_epoch(0, train=False, batches=1, save_reconstruction=False)
    assert np.isfinite(ts["test/loss"])
    t.save_model("final", directory=str(tmp_path))
    sd = torch.load(os.path.join(str(tmp_path), "final_params.pkl"), map_location="cpu")
    assert len(sd) == 85 and all(k.startswith("module.") for k in sd)


def test_cuda_graph_train_step_is_bit_identical_to_eager():
    """OP3Trainer.use_cuda_graph: forward+backward captured once per (shapes, schedule) and replayed.  With
    deterministic sampling (noise = 0 in both paths) five graph steps must reproduce five eager steps bit for bit,
    including new input batches copied into the graph's static buffers and weights refreshed inside the graph."""
    import op3_b200
    from op3_b200.torch.op3_modules.op3_trainer import OP3Trainer, TrainingScheduler
    B, K, T = 4, 3, 3
    sched, lw = np.array([0., 0., 1., 1., 0.]), np.arange(1, 6)
    batches = [(orc.synth_frames(B, T, seed=20 + i).cuda(), orc.synth_actions(B, T, seed=40 + i).cuda()) for i in range(3)]

    def run(graph):
        m, _ = build_model(4, K, seed=0)
        m.deterministic_sampling = True
        tr = OP3Trainer(None, None, m, TrainingScheduler(2, "curriculum", T, T), B, lr=3e-4)
        tr.use_cuda_graph = graph
        vals = []
        for i in range(6):
            fr, ac = batches[i % 3]
            vals.append([float(v) for v in tr.train_step(fr, ac, sched, lw)])
        if graph:
            assert any("graph" in e for e in tr._graphs.values()), "the step was never captured"
        return vals, {k: v.detach().clone() for k, v in m.named_parameters()}

    op3_b200.set_precision("f16x3")
    try:
        v_eager, p_eager = run(False)
        v_graph, p_graph = run(True)
    finally:
        op3_b200.set_precision("fp32")
    assert v_graph == v_eager, "losses differ:\n%r\n%r" % (v_graph, v_eager)
    for k in p_eager:
        assert torch.equal(p_eager[k], p_graph[k]), k


def test_wgrad_lane_is_bit_identical_to_serial():
    """Engine.wgrad_overlap (op3_set_wgrad_overlap / op3_wgrad_join): the conv weight-gradient kernels of the backward pass run on
    the library's side stream next to the dgrad chain.  Same accumulation order per parameter, so three optimizer steps must
    give the same losses and weights bit for bit as the serial order -- any operand overwritten early would show up here."""
    import op3_b200
    from op3_b200.engine import Engine
    from op3_b200.torch.op3_modules.op3_trainer import OP3Trainer, TrainingScheduler
    B, K, T = 16, 4, 3
    sched, lw = np.array([0., 0., 1., 0., 1., 0.]), np.arange(1, 7)
    batches = [(orc.synth_frames(B, T, seed=70 + i).cuda(), orc.synth_actions(B, T, seed=90 + i).cuda()) for i in range(3)]

    def run(overlap):
        m, _ = build_model(4, K, seed=0)
        m.deterministic_sampling = True
        m.engine.wgrad_overlap = overlap
        tr = OP3Trainer(None, None, m, TrainingScheduler(2, "curriculum", T, T), B, lr=3e-4)
        vals = [[float(v) for v in tr.train_step(*batches[i], sched, lw)] for i in range(3)]
        return vals, {k: v.detach().clone() for k, v in m.named_parameters()}

    assert Engine.wgrad_overlap is True, "the lane is the default"
    for prec in ("f16x3", "fp32"):
        op3_b200.set_precision(prec)
        try:
            v_ser, p_ser = run(False)
            v_lane, p_lane = run(True)
        finally:
            op3_b200.set_precision("fp32")
        assert v_lane == v_ser, "%s: losses differ:\n%r\n%r" % (prec, v_lane, v_ser)
        for k in p_ser:
            assert torch.equal(p_ser[k], p_lane[k]), "%s: %s" % (prec, k)
    assert op3_b200._lib.load().op3_wgrad_join(torch.cuda.current_stream().cuda_stream) == 0      # nothing pending: a no-op


def test_train_epoch_lagged_readback_equals_blocking(tmp_path):
    """OP3Trainer.stat_lag: train_epoch copies every step's scalars to pinned memory and reads them one step later (the GPU no
    longer idles while the host prepares the next step).  Same epoch statistics as the blocking read of the reference."""
    from torch.utils.data.dataset import TensorDataset
    from op3_b200.torch.data_management.dataset import BlocksDataset
    from op3_b200.torch.op3_modules.op3_trainer import OP3Trainer, TrainingScheduler, ScalarReadback
    torch.manual_seed(0)
    feats = torch.randint(0, 256, (12, 5, 3, 64, 64)).float()
    actions = torch.randn(12, 5, 6)

    def run(lag):
        train = BlocksDataset(TensorDataset(feats, actions), batchsize=4, shuffle=False)
        train.action_dim = 4
        m, _ = build_model(4, 3, seed=0)
        m.deterministic_sampling = True
        t = OP3Trainer(train, train, m, TrainingScheduler(seed_steps=2, schedule_type="curriculum", max_T=5, T=3), batch_size=4, lr=3e-4)
        t.stat_lag = lag
        return t.train_epoch(1)

    a, b = run(0), run(1)
    for k in ("train/loss", "train/Log Prob", "train/KL", "train/mse"):
        assert a[k] == b[k] and np.isfinite(a[k]), (k, a[k], b[k])
    rb = ScalarReadback(torch.device("cuda", 0), 2, lag=1, slots=3)          # ring wrap-around, order, drain
    rows = []
    for i in range(7):
        rows += rb.push([torch.tensor(float(i), device="cuda"), torch.tensor(float(-i), device="cuda")])
        assert len(rows) == i
    rows += rb.drain()
    assert rows == [[float(i), float(-i)] for i in range(7)]
