"""Generate golden vectors by running the UNMODIFIED reference (imported read-only from /root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Weights come from ``oracle.op3_oracle.make_params`` (deterministic from shapes+seed) and are loaded into the
reference model with ``load_state_dict(strict=True)``; sampling noise is injected by replacing
``op3.torch.pytorch_util.randn`` with a queue of pre-generated tensors.  Outputs are
written as small ``.npz`` files next to this script.  Large tensors (parameter gradients) are stored as
per-tensor (L2 norm, sum, 32 sampled entries) so fixtures stay small.
"""
import os
import sys
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

for mod in ["matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.patches", "matplotlib.cm",
            "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.mplot3d", "imageio", "mujoco_py", "h5py", "pathos",
            "pathos.pools", "doodad"]:
    sys.modules.setdefault(mod, MagicMock())

import op3.torch.pytorch_util as ptu                      # noqa: E402
import op3.torch.op3_modules.op3_model as ref_model       # noqa: E402
from op3.exp_variants.variants import stack_variant, pickplace_variant   # noqa: E402
import exps.pickplace_exps.mpc_pickplace as ref_mpc       # noqa: E402

from oracle import op3_oracle as orc                      # noqa: E402

torch.set_num_threads(os.cpu_count())


class NoiseQueue:
    def __init__(self):
        self.q = []

    def load(self, tensors):
        self.q = list(tensors)

    def __call__(self, *size, **kw):
        if not self.q:                       # model constructors also draw (op3_model.py:92-94); overwritten later
            return torch.randn(*size)
        t = self.q.pop(0)
        assert tuple(t.shape) == tuple(size), (t.shape, size)
        return t.clone()


NOISE = NoiseQueue()
ptu.randn = NOISE


def build_reference(variant, K, action_dim, seed, dtype=torch.float32):
    a = dict(variant["op3_args"])
    a["K"] = K
    m = ref_model.create_model_v2(a, a["det_repsize"], a["sto_repsize"], action_dim)
    p = orc.make_params(action_dim, seed=seed, dtype=dtype)
    m.load_state_dict(p, strict=True)
    return m, p


synth_frames, synth_actions, sample_idx = orc.synth_frames, orc.synth_actions, orc.sample_idx


def summarize(prefix, t, out):
    t = t.detach().double().reshape(-1)
    idx = sample_idx(t.numel())
    out[prefix + "/norm"] = np.float64(t.norm().item())
    out[prefix + "/sum"] = np.float64(t.sum().item())
    out[prefix + "/idx"] = idx.numpy()
    out[prefix + "/val"] = t[idx].numpy()


def train_case(name, variant, K, action_dim, B, T, schedule, full_masks):
    m, p = build_reference(variant, K, action_dim, seed=0)
    frames = synth_frames(B, T)
    actions = synth_actions(B, T) if action_dim else None
    schedule = np.array(schedule, dtype=np.float64)
    lw = np.arange(1, len(schedule) + 1)
    NOISE.load(orc.make_noise(B, K, 64, len(schedule) + 1))
    out = m.forward(frames, actions, None, schedule, lw)
    out[3].backward()
    res = {"schedule": schedule, "loss_schedule": lw, "B": B, "K": K, "T": T, "action_dim": action_dim,
           "total_loss": out[3].item(), "kle_loss": out[4].item(), "clog_prob": out[5].item(), "mse": out[6].item()}
    if full_masks:
        res["masks_list"] = out[1].detach().numpy()
        res["final_recon"] = out[2].detach().numpy()
    summarize("masks_list", out[1], res)
    summarize("colors_list", out[0], res)
    summarize("final_recon_s", out[2], res)
    for k in ["lambdas1", "lambdas2", "samples", "deter_state"]:
        summarize("state/" + k, out[7]["post"][k], res)
    summarize("state/h", out[7]["post"]["extra_info"][0], res)
    summarize("state/c", out[7]["post"]["extra_info"][1], res)
    for k, v in m.named_parameters():
        summarize("grad/" + k, v.grad if v.grad is not None else torch.zeros_like(v), res)
    total_norm = torch.nn.utils.clip_grad_norm_(list(m.parameters()), 5.0)
    res["grad_total_norm"] = float(total_norm)
    opt = torch.optim.Adam(list(m.parameters()), lr=3e-4)
    opt.step()
    for k, v in m.named_parameters():
        summarize("adam/" + k, v.detach() - p[k], res)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
    print(name, "loss", res["total_loss"], "gnorm", res["grad_total_norm"])


def noshare_case(name, variant, nets, K, action_dim, B, T, schedule):
    """(f3) the *_No_Sharing registry entries (decoder_network.py:92, refinement_network.py:235, physics_network.py:192): one
    training step of the unmodified reference with per-slot weight sets from orc.make_params_no_share."""
    a = dict(variant["op3_args"])
    a["K"] = K
    types = {"decode_net": ("decoder_model_type", "reg_no_share"), "refinement_net": ("refinement_model_type", "size_dependent_conv_no_share"),
             "dynamics_net": ("dynamics_model_type", "reg_ac32_no_share")}
    for n in nets:
        a[types[n][0]] = types[n][1]
    m = ref_model.create_model_v2(a, a["det_repsize"], a["sto_repsize"], action_dim)
    p = orc.make_params_no_share(action_dim, K, nets=nets, seed=0)
    m.load_state_dict(p, strict=True)
    frames = synth_frames(B, T)
    actions = synth_actions(B, T) if action_dim else None
    schedule = np.array(schedule, dtype=np.float64)
    lw = np.arange(1, len(schedule) + 1)
    NOISE.load(orc.make_noise(B, K, 64, len(schedule) + 1))
    out = m.forward(frames, actions, None, schedule, lw)
    out[3].backward()
    res = {"schedule": schedule, "loss_schedule": lw, "B": B, "K": K, "T": T, "action_dim": action_dim, "nets": np.array(nets),
           "total_loss": out[3].item(), "kle_loss": out[4].item(), "clog_prob": out[5].item(), "mse": out[6].item(),
           "masks_list": out[1].detach().numpy(), "final_recon": out[2].detach().numpy()}
    summarize("colors_list", out[0], res)
    for k in ["lambdas1", "lambdas2", "samples", "deter_state"]:
        summarize("state/" + k, out[7]["post"][k], res)
    summarize("state/h", out[7]["post"]["extra_info"][0], res)
    summarize("state/c", out[7]["post"]["extra_info"][1], res)
    for k, v in m.named_parameters():
        summarize("grad/" + k, v.grad if v.grad is not None else torch.zeros_like(v), res)
    res["grad_total_norm"] = float(torch.nn.utils.clip_grad_norm_(list(m.parameters()), 5.0))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **res)
    print(name, "loss", res["total_loss"], "gnorm", res["grad_total_norm"])


def noshare_cases():
    noshare_case("noshare_all", pickplace_variant, orc.NETS, K=3, action_dim=4, B=2, T=3, schedule=[0, 0, 1, 0, 1])
    noshare_case("noshare_dec_ref", stack_variant, ("decode_net", "refinement_net"), K=3, action_dim=0, B=3, T=2, schedule=[0, 0, 1])
    noshare_case("noshare_dyn", pickplace_variant, ("dynamics_net",), K=4, action_dim=4, B=2, T=2, schedule=[0, 1, 0])
    # the other two registry entries of refinement_network.py:15-62 cannot even be constructed in the reference: the shape
    # probe of RefinementNetwork.__init__ (:150-158) runs AvgPool2d(input_width) on a 52x52 (large_reg: paddings 0) / 1x1
    # (large_sequence: stride 2) feature map and raises
    errs = {}
    for key in ("large_reg", "large_sequence"):
        a = dict(pickplace_variant["op3_args"])
        a["K"] = 2
        a["refinement_model_type"] = key
        try:
            m = ref_model.create_model_v2(a, a["det_repsize"], a["sto_repsize"], 4)       # already fails in __init__ (:157)
            m.forward(synth_frames(1, 2), synth_actions(1, 2), None, np.array([0.0, 1.0]), np.array([1.0, 2.0]))
            errs[key] = "ran"
        except RuntimeError as e:
            errs[key] = "RuntimeError: " + str(e)[:120]
    np.savez_compressed(os.path.join(HERE, "registry_unrunnable.npz"), **{k: np.array(v) for k, v in errs.items()})
    print(errs)


def module_case():
    """Per-module forward goldens (decoder, refinement net, dynamics net) + refine-input pieces."""
    m, p = build_reference(pickplace_variant, 3, 4, seed=3)
    g = torch.Generator().manual_seed(21)
    res = {}
    z = torch.randn(6, 128, generator=g) * 0.7
    logits, colors = m.decode_net(z)
    res["dec/z"] = z.numpy()
    summarize("dec/colors", colors, res)
    summarize("dec/logits", logits, res)
    a = torch.randn(6, 15, 64, 64, generator=g)
    h = torch.randn(6, 128, generator=g) * 0.3
    c = torch.randn(6, 128, generator=g) * 0.3
    extra = torch.randn(6, 320, generator=g)
    l1, l2, h2, c2 = m.refinement_net(a, h, c, extra)
    res["ref/seed"] = 21
    for k, v in [("l1", l1), ("l2", l2), ("h2", h2[0]), ("c2", c2[0])]:
        res["ref/" + k] = v.detach().numpy()
    zz = torch.randn(6, 128, generator=g) * 0.7
    acts = synth_actions(2, 1, seed=4)[:, 0].unsqueeze(1).repeat(1, 3, 1).reshape(6, 6)
    d, l1, l2 = m.dynamics_net(zz, acts)
    res["dyn/z"] = zz.numpy()
    res["dyn/actions"] = acts.numpy()
    for k, v in [("deter", d), ("l1", l1), ("l2", l2)]:
        res["dyn/" + k] = v.detach().numpy()
    sa, ia, dl = m.dynamics_net.get_all_attention_values(zz, acts[:, [0, 1, 3, 4]], 3)
    res["dyn/act_att"] = sa.detach().numpy()
    res["dyn/pair_att"] = ia.detach().numpy()
    res["dyn/deltas"] = dl.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "modules.npz"), **res)
    print("modules ok")


COST_OPTION_CASES = [("latent", "mse", "raw", "sum"), ("latent", "mse", "raw", "min"), ("subimage", "psuedo_intersect", "raw", "sum"),
                     ("subimage", "psuedo_intersect", "raw", "min"), ("subimage", "mse", "negative_exp", "sum"),
                     ("final_recon", "mse", "negative_exp", "sum"), ("latent", "mse", "negative_exp", "min")]


def mpc_case():
    """pickplace-style MPC rollout: B=1 state from [0,0,1,0] on real-shaped synthetic frames, then Na candidate
    actions through schedule [1] in chunks of 10 (op3_model.py:569-642) and Cost ranking (mpc_pickplace.py:50)."""
    K, Na = 4, 24
    m, p = build_reference(pickplace_variant, K, 4, seed=7)
    m.eval_mode = True
    frames = synth_frames(1, 2, seed=31)
    acts = synth_actions(1, 2, seed=32)[:, :, [0, 1, 3, 4]]
    sched = np.array([0, 0, 1, 0], dtype=np.float64)
    NOISE.load(orc.make_noise(1, K, 64, len(sched) + 1, seed=40))
    info = m.batch_internal_inference(frames, acts, None, sched)
    state = info["state"]
    cand = synth_actions(Na, 1, seed=33)[:, :, [0, 1, 3, 4]]
    full_noise = orc.make_noise(Na, K, 64, 1, seed=50)[0]
    NOISE.load([full_noise[i:i + 10] for i in range(0, Na, 10)])
    pred = m.batch_internal_inference(None, cand, state, np.array([1.0]))
    goal_image = synth_frames(1, 1, seed=34)[:, 0]
    res = {"K": K, "Na": Na}
    summarize("state/samples", state["post"]["samples"], res)
    summarize("pred/final_recon", pred["final_recon"], res)
    summarize("pred/sub_images", pred["sub_images"], res)
    for core, agg in [("final_recon", "sum"), ("subimage", "min"), ("subimage", "sum")]:
        cost = ref_mpc.Cost(None, core_type=core, compare_func="mse", post_process="raw", aggregate=agg)
        sc, order, gidx = cost.get_action_rankings(info["state"]["post"]["samples"][0], info["sub_images"][0],
                                                   goal_image, pred["state"]["post"]["samples"], pred["sub_images"],
                                                   pred["final_recon"], plot_actions=0)
        res["cost/%s_%s/sorted" % (core, agg)] = np.asarray(sc)
        res["cost/%s_%s/order" % (core, agg)] = np.asarray(order)
        res["cost/%s_%s/gidx" % (core, agg)] = np.asarray(gidx)
    # the remaining Cost options (mpc_stack.py:60-100): latent distances, 'psuedo_intersect', 'negative_exp'
    for core, cmp_, post, agg in COST_OPTION_CASES:
        cost = ref_mpc.Cost(None, core_type=core, compare_func=cmp_, post_process=post, aggregate=agg)
        sc, order, gidx = cost.get_action_rankings(info["state"]["post"]["samples"][0], info["sub_images"][0],
                                                   goal_image, pred["state"]["post"]["samples"], pred["sub_images"],
                                                   pred["final_recon"], plot_actions=0)
        key = "costx/%s_%s_%s_%s/" % (core, cmp_, post, agg)
        res[key + "sorted"], res[key + "order"], res[key + "gidx"] = np.asarray(sc), np.asarray(order), np.asarray(gidx)
    np.savez_compressed(os.path.join(HERE, "mpc.npz"), **res)
    print("mpc ok", res["cost/final_recon_sum/order"][:5])


def mpc_large_case(Na=4096):
    """BASELINE.json's MPC size: the reference itself (chunks of 10, op3_model.py:569-642) rolls 4096 candidate actions
    out of the B=1 state of mpc_case and ranks them with Cost(final_recon, mse, sum) (mpc_pickplace.py:50-236)."""
    K = 4
    m, p = build_reference(pickplace_variant, K, 4, seed=7)
    m.eval_mode = True
    frames = synth_frames(1, 2, seed=31)
    acts = synth_actions(1, 2, seed=32)[:, :, [0, 1, 3, 4]]
    sched = np.array([0, 0, 1, 0], dtype=np.float64)
    NOISE.load(orc.make_noise(1, K, 64, len(sched) + 1, seed=40))
    info = m.batch_internal_inference(frames, acts, None, sched)
    cand = synth_actions(Na, 1, seed=33)[:, :, [0, 1, 3, 4]]
    full_noise = orc.make_noise(Na, K, 64, 1, seed=50)[0]
    NOISE.load([full_noise[i:i + 10] for i in range(0, Na, 10)])
    pred = m.batch_internal_inference(None, cand, info["state"], np.array([1.0]))
    goal_image = synth_frames(1, 1, seed=34)[:, 0]
    cost = ref_mpc.Cost(None, core_type="final_recon", compare_func="mse", post_process="raw", aggregate="sum")
    sc, order, gidx = cost.get_action_rankings(info["state"]["post"]["samples"][0], info["sub_images"][0], goal_image,
                                               pred["state"]["post"]["samples"], pred["sub_images"], pred["final_recon"],
                                               plot_actions=0)
    np.savez_compressed(os.path.join(HERE, "mpc_4096.npz"), Na=Na, K=K, sorted=np.asarray(sc, dtype=np.float32),
                        order=np.asarray(order, dtype=np.int32))
    print("mpc_4096 ok", np.asarray(order)[:5], np.asarray(sc)[:3])


def activation_case():
    """OP3Model.get_all_activation_values / get_activation_values (op3_model.py:645-698) for 23 probe actions on the
    B=1 state of mpc_case (pickplace latent-action MPC, mpc_pickplace.py:663-703)."""
    K, Np = 4, 23
    m, p = build_reference(pickplace_variant, K, 4, seed=7)
    m.eval_mode = True
    frames = synth_frames(1, 2, seed=31)
    acts = synth_actions(1, 2, seed=32)[:, :, [0, 1, 3, 4]]
    NOISE.load(orc.make_noise(1, K, 64, 5, seed=40))
    info = m.batch_internal_inference(frames, acts, None, np.array([0, 0, 1, 0], dtype=np.float64))
    probe = synth_actions(Np, 1, seed=35)[:, 0][:, [0, 1, 3, 4]]
    sa, ia, dv, ld = m.get_all_activation_values(info["state"], probe)
    sa1 = m.get_activation_values(info["state"], probe)
    np.savez_compressed(os.path.join(HERE, "activations.npz"), K=K, Np=Np, state_action=sa.numpy(), interaction=ia.numpy(),
                        deltas=dv.numpy(), lambda_deltas=ld.numpy(), state_action_only=sa1.detach().numpy())
    print("activations ok", sa.shape, ia.shape)


def chunked_noise(full, chunk):
    """Reference draw order of batch_internal_inference (op3_model.py:569-642): per chunk of `chunk` samples one draw
    for the initial state (if it is created) and one per schedule step.  `full` = list of (Na,K,Rs) tensors."""
    Na = full[0].shape[0]
    return [f[i:i + chunk] for i in range(0, Na, chunk) for f in full]


def mpc_stack_case(Na=4096):
    """BASELINE.json config 5, stack flavour (mpc_stack.py:249-258): Na candidate AFTER-images (B,T=1,3,D,D), no
    actions, no initial state, schedule [0,0,0,0,1] (K=7, action_dim 0), ranked with Cost(subimage, mse, raw, min)
    (mpc_stack.py:553-558) against the goal latents of a goal frame refined 5 times (mpc_stack.py:403-423)."""
    K = 7
    m, p = build_reference(stack_variant, K, 0, seed=7)
    m.eval_mode = True
    goal_frames = synth_frames(1, 1, seed=61)
    NOISE.load(orc.make_noise(1, K, 64, 6, seed=62))
    goal = m.batch_internal_inference(goal_frames, None, None, np.zeros(5))
    obs = synth_frames(Na, 1, seed=63)
    sched = np.array([0, 0, 0, 0, 1], dtype=np.float64)
    full = orc.make_noise(Na, K, 64, len(sched) + 1, seed=64)
    NOISE.load(chunked_noise(full, 10))
    pred = m.batch_internal_inference(obs, None, None, sched)
    import exps.stack_exps.mpc_stack as ref_stack
    cost = ref_stack.Cost(None, core_type="subimage", compare_func="mse", post_process="raw", aggregate="min")
    sc, order, gidx = cost.get_action_rankings(goal["state"]["post"]["samples"][0], goal["sub_images"][0], goal_frames[:, 0],
                                               pred["state"]["post"]["samples"], pred["sub_images"], pred["final_recon"],
                                               plot_actions=0)
    res = {"Na": Na, "K": K, "sorted": np.asarray(sc, dtype=np.float32), "order": np.asarray(order, dtype=np.int32),
           "gidx": np.asarray(gidx, dtype=np.int32)}
    summarize("goal/sub_images", goal["sub_images"], res)
    summarize("pred/final_recon", pred["final_recon"], res)
    summarize("pred/sub_images", pred["sub_images"], res)
    summarize("pred/samples", pred["state"]["post"]["samples"], res)
    np.savez_compressed(os.path.join(HERE, "mpc_stack_%d.npz" % Na), **res)
    print("mpc_stack ok", np.asarray(order)[:5], np.asarray(sc)[:3])


def pretrained_case():
    """REPORTED (not gated): the three shipped checkpoints on the shipped real frames
    (data/goals), one training step each in fp32 AND fp64 of the reference.  Weights + inputs + outputs are committed
    so that the GPU box (no /root/reference) can replay them (tests/test_gpu_pretrained.py)."""
    import pickle
    REF = "/root/reference"
    cases = [
        ("stack", "exps/stack_exps/saved_models/s64d64_v1_params.pkl", stack_variant, 7, 0, [0, 0, 0, 0, 1]),
        ("pickplace", "exps/pickplace_exps/saved_models/curriculum_aws_params.pkl", pickplace_variant, 4, 4,
         [0, 0, 0, 0, 1, 1, 1, 1, 0]),
        ("cloth", "exps/realworld_exps/saved_models/cloth_reg_params.pkl", pickplace_variant, 4, 4, [0, 0, 0, 0, 1, 1, 0]),
    ]
    # real frames: stack before/after pairs, pickplace 5-frame trajectories with 6-dim actions
    pairs = []
    for name in ["bridge", "air-cube", "double-bridge", "air-triangle"]:
        d = pickle.load(open(os.path.join(REF, "data/goals/stack_goals/manual_constructions", name, "0.p"), "rb"))
        pairs.append(np.asarray(d["images"], dtype=np.uint8))                      # (2,64,64,3)
    stack_frames = np.stack(pairs)                                                 # (4,2,64,64,3)
    gd = pickle.load(open(os.path.join(REF, "data/goals/pickplace_goals/objects_seed_2/goal_data.pkl"), "rb"))
    pp_frames = np.asarray(gd["frames"][:4], dtype=np.uint8)                       # (4,5,64,64,3)
    pp_actions = np.asarray(gd["actions"][:4], dtype=np.float32)                   # (4,4,6)
    for name, ckpt, variant, K, A, sched in cases:
        sd = torch.load(os.path.join(REF, ckpt), map_location="cpu")
        sd = {k[len("module."):] if k.startswith("module.") else k: v.float() for k, v in sd.items()}
        if name == "stack":
            fr_u8, acts = stack_frames, None
        else:
            fr_u8 = pp_frames
            acts = torch.from_numpy(np.concatenate([pp_actions, pp_actions[:, -1:]], 1))   # (4,5,6): T actions
        B, T = fr_u8.shape[:2]
        frames = torch.from_numpy(fr_u8).permute(0, 1, 4, 2, 3).float() / 255.0
        schedule = np.array(sched, dtype=np.float64)
        lw = np.arange(1, len(sched) + 1)
        noise = orc.make_noise(B, K, 64, len(sched) + 1, seed=70)
        res = {"schedule": schedule, "K": K, "action_dim": A, "frames_u8": fr_u8}
        if acts is not None:
            res["actions"] = acts.numpy()
        for k, v in sd.items():
            res["w/" + k] = v.numpy()
        for tag, dt in [("f32", torch.float32), ("f64", torch.float64)]:
            torch.set_default_dtype(dt)          # ptu.zeros / torch.zeros inside the reference follow the default
            a = dict(variant["op3_args"]); a["K"] = K
            m = ref_model.create_model_v2(a, a["det_repsize"], a["sto_repsize"], A)
            m.load_state_dict(sd, strict=True)
            m = m.to(dt)
            m.decode_net.broadcast_net.coords = m.decode_net.broadcast_net.coords.to(dt)
            m.refinement_net.coords = m.refinement_net.coords.to(dt)
            NOISE.load([n.to(dt) for n in noise])
            out = m.forward(frames.to(dt), None if acts is None else acts.to(dt), None, schedule, lw)
            out[3].backward()
            res[tag + "/total_loss"] = out[3].item(); res[tag + "/kle_loss"] = out[4].item()
            res[tag + "/clog_prob"] = out[5].item(); res[tag + "/mse"] = out[6].item()
            res[tag + "/masks_last"] = out[1][:, -1].detach().float().numpy().astype(np.float16)
            res[tag + "/final_recon"] = out[2].detach().float().numpy()
            summarize(tag + "/masks_list", out[1], res)
            for k, v in m.named_parameters():
                summarize(tag + "/grad/" + k, v.grad if v.grad is not None else torch.zeros_like(v), res)
            res[tag + "/grad_total_norm"] = float(torch.sqrt(sum((v.grad.double() ** 2).sum() for v in m.parameters()
                                                                 if v.grad is not None)))
            print("pretrained", name, tag, "loss", res[tag + "/total_loss"], "mse", res[tag + "/mse"])
        torch.set_default_dtype(torch.float32)
        np.savez_compressed(os.path.join(HERE, "pretrained_%s.npz" % name), **res)


def schedule_case():
    """TrainingScheduler outputs (op3_trainer.py:36-111) for every schedule type under a fixed numpy seed."""
    from op3.torch.op3_modules.op3_trainer import TrainingScheduler as Ref
    cases = []
    for kind in ["single_step_physics", "multi_step_physics", "curriculum", "rprp_curriculum", "static_iodine", "rprp",
                 "next_step", "random_alternating"]:
        for epoch in [0, 49, 100, 101, 135, 160, 400]:
            for is_train in [True, False]:
                for max_T in [None, 3, 5]:
                    np.random.seed(3)
                    r = Ref(4, kind, max_T, 5)
                    cases.append((kind, epoch, is_train, max_T, r.get_schedule(epoch, is_train), r.should_save_model(epoch)))
    arr = np.empty(len(cases), dtype=object)
    for i, c in enumerate(cases):
        arr[i] = c
    np.savez_compressed(os.path.join(HERE, "schedules.npz"), cases=arr)
    print("schedules", len(cases))


if __name__ == "__main__":
    only = sys.argv[1:]
    if only:                       # e.g. `make_golden.py mpc_stack pretrained`: regenerate selected (slow) cases only
        for name in only:
            {"mpc_stack": mpc_stack_case, "pretrained": pretrained_case, "mpc_large": mpc_large_case,
             "mpc_stack_small": lambda: mpc_stack_case(Na=40), "activations": activation_case, "noshare": noshare_cases, "mpc": mpc_case}[name]()
        sys.exit(0)
    schedule_case()
    train_case("stack_small", stack_variant, K=3, action_dim=0, B=2, T=2, schedule=[0, 0, 1], full_masks=True)
    train_case("pickplace_small", pickplace_variant, K=3, action_dim=4, B=2, T=3, schedule=[0, 0, 1, 1, 0],
               full_masks=False)
    train_case("stack_c1", stack_variant, K=7, action_dim=0, B=8, T=2, schedule=[0, 0, 0, 0, 1], full_masks=False)
    train_case("pickplace_nextstep", pickplace_variant, K=3, action_dim=4, B=2, T=4, schedule=[0, 2, 2, 1, 0],
               full_masks=False)
    module_case()
    mpc_case()
    mpc_large_case()
    activation_case()
    noshare_cases()
    pretrained_case()
    mpc_stack_case(Na=40)
    mpc_stack_case()               # ~30-60 min of CPU: 4096 candidates x (5 decodes + 4 refinements), K=7
