"""CPU oracle for the OP3 iterative amortized-inference step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``op3_b200/`` may import this module.  It is used by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` as the *checker* for the sm_100a kernels, never as a product path.

It is a functional (state_dict-keyed, no nn.Module) restatement of the reference algorithm in
plain PyTorch-on-CPU arithmetic.  Every function cites the reference file:line it restates
(paths relative to the reference checkout, ``op3/torch/...``).

Parity pinning: the reference ships no tests or golden vectors.  The oracle is
therefore pinned against outputs of the *unmodified reference itself*, imported in the build
container by ``tests/golden/make_golden.py``; the resulting vectors are committed under
``tests/golden/`` and ``tests/test_oracle_golden.py`` checks this file against them on CPU.

The third-party arithmetic of the reference is PyTorch itself (conv2d, LSTM, Linear, softmax,
MultivariateNormal KL; pinned pytorch=1.0.0 in environment/linux-gpu-env.yml:12).  The closed forms
used below (diagonal-Gaussian KL, LSTM cell, closed-form refinement gradients) restate the
published definitions of those ops.
"""
from collections import OrderedDict
import math

import numpy as np
import torch
import torch.nn.functional as F

SIGMA = 0.1          # op3_model.py:79
LSTM_SIZE = 128      # refinement_network.py:41 (size_dependent_conv)
ACTION_ENC = 32      # physics_network.py:18 (reg_ac32)


# ----------------------------------------------------------------------------------------------
# Parameters (Appendix A of SURVEY.md; key names are the reference state_dict keys)
# ----------------------------------------------------------------------------------------------
def param_shapes(action_dim, det=64, sto=64):
    """Ordered (key, shape) list of the OP3 state_dict (op3_model.py:91-94, decoder_network.py:10-23,
    refinement_network.py:31-46,134-179, physics_network.py:56-87)."""
    R = det + sto
    s = []
    if det:
        s.append(("inital_deter_state", (det,)))      # sic, op3_model.py:92
    s += [("initial_lambdas1", (sto,)), ("initial_lambdas2", (sto,))]
    s += [("refinement_net.lstm.weight_ih_l0", (4 * LSTM_SIZE, sto * 5 + 128)),
          ("refinement_net.lstm.weight_hh_l0", (4 * LSTM_SIZE, LSTM_SIZE)),
          ("refinement_net.lstm.bias_ih_l0", (4 * LSTM_SIZE,)),
          ("refinement_net.lstm.bias_hh_l0", (4 * LSTM_SIZE,))]
    cin = 17
    for i, co in enumerate([32, 32, sto]):
        s += [("refinement_net.conv_layers.%d.weight" % i, (co, cin, 5, 5)),
              ("refinement_net.conv_layers.%d.bias" % i, (co,))]
        cin = co
    s += [("refinement_net.fc_layers.0.weight", (128, sto)), ("refinement_net.fc_layers.0.bias", (128,)),
          ("refinement_net.last_fc.weight", (sto, LSTM_SIZE)), ("refinement_net.last_fc.bias", (sto,)),
          ("refinement_net.last_fc2.weight", (sto, LSTM_SIZE)), ("refinement_net.last_fc2.bias", (sto,))]

    def mlp(name, hidden, out, inp):
        return [("dynamics_net.%s.fcs.0.weight" % name, (hidden, inp)),
                ("dynamics_net.%s.fcs.0.bias" % name, (hidden,)),
                ("dynamics_net.%s.last_fc.weight" % name, (out, hidden)),
                ("dynamics_net.%s.last_fc.bias" % name, (out,))]
    s += mlp("inertia_encoder", R, R, R)
    if action_dim > 0:
        s += mlp("action_encoder", R, ACTION_ENC, action_dim)
        s += mlp("action_effect_network", R, R, ACTION_ENC + R)
        s += mlp("action_attention_network", R, 1, ACTION_ENC + R)
    s += mlp("pairwise_encoder_network", 2 * R, R, 2 * R)
    s += mlp("interaction_effect_network", R, ACTION_ENC, R)
    s += mlp("interaction_attention_network", R, 1, R)
    s += mlp("final_merge_network", R, R, ACTION_ENC + R)
    if det:
        s += mlp("det_output", R, det, R)
    s += mlp("lambdas1_output", R, sto, R)
    s += mlp("lambdas2_output", R, sto, R)
    cin = R + 2
    for i, co in enumerate([32, 32, 32, 32, 4]):
        s += [("decode_net.broadcast_net.conv_layers.%d.weight" % i, (co, cin, 5, 5)),
              ("decode_net.broadcast_net.conv_layers.%d.bias" % i, (co,))]
        cin = co
    for i, c in enumerate([1, 1, 1, 3]):
        s += [("layer_norms_2d.%d.gamma" % i, (c,)), ("layer_norms_2d.%d.beta" % i, (c,))]
    for i in range(2):
        s += [("layer_norms_1d.%d.scale_param" % i, (sto,)), ("layer_norms_1d.%d.center_param" % i, (sto,))]
    return s


def make_params(action_dim, det=64, sto=64, seed=0, dtype=torch.float32, generic=True):
    """Deterministic synthetic weights, reproducible from (shapes, seed) alone so that fixtures do not
    have to carry 3.5 MB of weights.  Distributions follow the reference initialisers in *scale*
    (xavier-uniform convs conv_networks.py:314, fan-in MLP networks.py:49-61, LayerNorm2D.gamma~U(0,1)
    modules.py:56) except that biases / tiny-init layers get O(fan-in) noise when ``generic`` so every
    term of every kernel is exercised (zero biases would hide bias-gradient bugs)."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for key, shape in param_shapes(action_dim, det, sto):
        n = int(np.prod(shape))
        u = torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1
        if key.endswith("gamma"):
            v = (u + 1) / 2
        elif key.endswith("scale_param"):
            v = 1 + 0.25 * u
        elif key.startswith("init") or key.startswith("inital"):
            v = u * math.sqrt(3.0 / shape[0])
        elif len(shape) == 4:
            fan_in, fan_out = shape[1] * 25, shape[0] * 25
            v = u * math.sqrt(6.0 / (fan_in + fan_out))
        elif len(shape) == 2:
            v = u / math.sqrt(shape[1])
        else:  # biases, LN beta / center
            v = u * (0.1 if generic else 0.0)
        p[key] = v.reshape(shape).to(dtype)
    return p


NETS = ("decode_net", "refinement_net", "dynamics_net")


def make_params_no_share(action_dim, K, nets=NETS, det=64, sto=64, seed=0, dtype=torch.float32):
    """Weights of a model whose sub-networks in ``nets`` are the *_No_Sharing variants: one weight set per slot under
    ``<net>.models.<k>.`` (decoder_network.py:113-131, refinement_network.py:268-287, physics_network.py:208-216); slot k's
    set comes from make_params(seed + 1000 (k + 1)), everything else from make_params(seed)."""
    base = make_params(action_dim, det, sto, seed, dtype)
    slots = [make_params(action_dim, det, sto, seed + 1000 * (k + 1), dtype) for k in range(K)]
    out = OrderedDict()
    for key, v in base.items():
        net = key.split(".")[0]
        if net in nets:
            for k in range(K):
                out["%s.models.%d.%s" % (net, k, key[len(net) + 1:])] = slots[k][key]
        else:
            out[key] = v
    return out


def _n_slots(p, net):
    """Number of per-slot weight sets of ``net`` in a state dict (0 = shared weights)."""
    pre = net + ".models."
    return len({key[len(pre):].split(".")[0] for key in p if key.startswith(pre)})


def _slot_view(p, net, k):
    """``p`` with the keys ``<net>.models.<k>.x`` exposed as ``<net>.x`` (the k-th sub-network as a shared-weight one)."""
    pre = "%s.models.%d." % (net, k)
    q = {key: v for key, v in p.items() if not key.startswith(net + ".models.")}
    q.update({net + "." + key[len(pre):]: v for key, v in p.items() if key.startswith(pre)})
    return q


def _ith(x, K, i):
    """rows (b, i) of a (bs*K, ...) tensor: the *_No_Sharing `_get_ith_input` (decoder_network.py:156-162)."""
    return None if x is None else x.reshape([-1, K] + list(x.shape[1:]))[:, i]


def make_noise(B, K, sto, n_draws, seed=5, dtype=torch.float32):
    """Pre-generated sampling noise, one (B,K,sto) tensor per ``ptu.randn`` draw
    (op3_model.py:151, called from :172, :253, :282).  Seeds 5+i as in SURVEY.md section 8d."""
    return [torch.randn(B, K, sto, generator=torch.Generator().manual_seed(seed + i),
                        dtype=torch.float64).to(dtype) for i in range(n_draws)]


def synth_frames(B, T, D=64, seed=1):
    """uint8 ~ U{0..255} frames / 255 as OP3Trainer.prepare_tensors does (op3_trainer.py:148-149)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (B, T, 3, D, D), generator=g, dtype=torch.uint8).float() / 255.0


def synth_actions(B, T, seed=2):
    """6-dim pick/place env actions (block_pick_and_place.py:56-58 bounds); the model uses cols [0,1,3,4]."""
    g = torch.Generator().manual_seed(seed)
    lo = torch.tensor([-2.5, 1.0, 0.59, -2.5, 1.0, 5.0])
    hi = torch.tensor([2.5, 4.0, 0.59, 2.5, 4.0, 5.0])
    return lo + (hi - lo) * torch.rand(B, T, 6, generator=g)


def sample_idx(n, count=32, seed=11):
    """Fixed pseudo-random positions used to store large tensors as small fixtures."""
    g = torch.Generator().manual_seed(seed + n % 9973)
    return torch.randint(0, n, (min(count, n),), generator=g)


# ----------------------------------------------------------------------------------------------
# Building blocks
# ----------------------------------------------------------------------------------------------
def coord_planes(D, dtype):
    """(1,2,D,D): channel 0 = x along width, channel 1 = y along height, linspace(-1,1,D)
    (conv_networks.py:321-324, refinement_network.py:179-182)."""
    lin = torch.linspace(-1, 1, D, dtype=torch.float64)
    xs = lin.view(1, D).expand(D, D)
    ys = lin.view(D, 1).expand(D, D)
    return torch.stack([xs, ys], 0).unsqueeze(0).to(dtype)


def elu(x):
    return torch.where(x > 0, x, torch.expm1(x))


def decoder(p, z, D=64):
    """Spatial-broadcast decoder (decoder_network.py:83-88, conv_networks.py:326-350).
    z (N,R) -> (N,4,D,D), channels [r,g,b,mask_logit]."""
    Kn = _n_slots(p, "decode_net")
    if Kn:      # DecoderNetwork_v2_No_Sharing.forward (decoder_network.py:144-154): slot-major torch.cat of the K outputs
        return torch.cat([decoder(_slot_view(p, "decode_net", k), _ith(z, Kn, k), D) for k in range(Kn)])
    N, R = z.shape
    h = torch.cat([z.reshape(N, R, 1, 1).expand(N, R, D, D), coord_planes(D, z.dtype).expand(N, 2, D, D)], 1)
    for i in range(5):
        pre = "decode_net.broadcast_net.conv_layers.%d." % i
        h = F.conv2d(h, p[pre + "weight"], p[pre + "bias"], padding=2)
        if i < 4:
            h = elu(h)
    return h


def softplus_std(x):
    """op3_model.py:130-132."""
    return torch.sqrt(torch.log(1 + torch.exp(torch.clamp(x, max=80.0))) + 1e-5)


def kl_post_prior(post_mu, post_sp, prior_mu, prior_sp):
    """KL(post || prior) for diagonal Gaussians; restates MultivariateNormal(scale_tril=diag) +
    kl_divergence (op3_model.py:137-145).  Returns (B,K)."""
    sq, sp = softplus_std(post_sp), softplus_std(prior_sp)
    t = torch.log(sp / sq) + (sq * sq + (post_mu - prior_mu) ** 2) / (2 * sp * sp) - 0.5
    return t.sum(-1)


def gaussian_prob(colors, target):
    """op3_model.py:123-126 with ch=3, sigma=0.1: exp(-S/(3*2*sigma^2)) / (sqrt(sigma^6)*248.05).
    colors (B,K,3,D,D), target (B,3,D,D) -> (B,K,D,D)."""
    S = ((colors - target.unsqueeze(1)) ** 2).sum(2)
    return torch.exp(-S / (3 * 2 * SIGMA ** 2)) / (math.sqrt(SIGMA ** 6) * 248.05)


def decode(p, state, D=64):
    """op3_model.py:300-308.  Returns colors (B,K,3,D,D), mask (B,K,1,D,D), mask_logits (B,K,1,D,D)."""
    post = state["post"]
    B, K = post["samples"].shape[:2]
    z = post["samples"] if post["deter_state"] is None else torch.cat([post["deter_state"], post["samples"]], 2)
    out = decoder(p, z.reshape(B * K, -1), D).view(B, K, 4, D, D)
    logits = out[:, :, 3:4]
    return out[:, :, :3], torch.softmax(logits, dim=1), logits


def get_loss(state, colors, mask, target, beta):
    """op3_model.py:313-327."""
    B = colors.shape[0]
    probs = gaussian_prob(colors, target)                       # (B,K,D,D)
    P = (mask.squeeze(2) * probs).sum(1)                        # (B,D,D)
    clog = -torch.log(P + 1e-12).sum() / B
    kl = kl_post_prior(state["post"]["lambdas1"], state["post"]["lambdas2"],
                       state["prior"]["lambdas1"], state["prior"]["lambdas2"]).sum() / B
    return probs, P, kl, clog, clog + beta * kl


def layernorm2d(x, gamma, beta, eps=1e-5):
    """modules.py:48-69: per-sample mean / unbiased std over (C,H,W), (x-mean)/(std+eps), per-channel affine."""
    n = x.shape[0]
    flat = x.reshape(n, -1)
    mean = flat.mean(1).view(n, 1, 1, 1)
    std = flat.std(1).view(n, 1, 1, 1)
    return (x - mean) / (std + eps) * gamma.view(1, -1, 1, 1) + beta.view(1, -1, 1, 1)


def layernorm1d(x, scale, center, eps=1e-6):
    """modules.py:19-46 with center=True, scale=True (op3_model.py:86)."""
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return (x - mean) / (std + eps) * scale + center


def refinement_net(p, a, h, c, extra):
    """refinement_network.py:189-219 ('size_dependent_conv').  a (N,15,D,D), h/c (N,128), extra (N,5*sto)."""
    Kn = _n_slots(p, "refinement_net")
    if Kn:      # RefinementNetwork_v2_No_Sharing.forward (refinement_network.py:291-307); h, c come back as rows k*bs + b
        outs = [refinement_net(_slot_view(p, "refinement_net", k), _ith(a, Kn, k), _ith(h, Kn, k), _ith(c, Kn, k),
                               _ith(extra, Kn, k)) for k in range(Kn)]
        return tuple(torch.cat([o[j] for o in outs]) for j in range(4))
    N, _, D, _ = a.shape
    x = torch.cat([a, coord_planes(D, a.dtype).expand(N, 2, D, D)], 1)
    for i in range(3):
        pre = "refinement_net.conv_layers.%d." % i
        x = elu(F.conv2d(x, p[pre + "weight"], p[pre + "bias"], padding=2))
    x = x.mean((2, 3))                                                     # AvgPool2d(D)
    x = elu(F.linear(x, p["refinement_net.fc_layers.0.weight"], p["refinement_net.fc_layers.0.bias"]))
    x = torch.cat([x, extra], 1)
    gates = (F.linear(x, p["refinement_net.lstm.weight_ih_l0"], p["refinement_net.lstm.bias_ih_l0"]) +
             F.linear(h, p["refinement_net.lstm.weight_hh_l0"], p["refinement_net.lstm.bias_hh_l0"]))
    gi, gf, gg, go = gates.chunk(4, 1)                                     # PyTorch LSTM gate order i,f,g,o
    c2 = torch.sigmoid(gf) * c + torch.sigmoid(gi) * torch.tanh(gg)
    h2 = torch.sigmoid(go) * torch.tanh(c2)
    lam1 = F.linear(h2, p["refinement_net.last_fc.weight"], p["refinement_net.last_fc.bias"])
    lam2 = F.linear(h2, p["refinement_net.last_fc2.weight"], p["refinement_net.last_fc2.bias"])
    return lam1, lam2, h2, c2


def _mlp(p, name, x, out_act):
    """networks.py:63-74 with one hidden layer, hidden_activation=ELU (physics_network.py:57-87)."""
    pre = "dynamics_net.%s." % name
    hdn = elu(F.linear(x, p[pre + "fcs.0.weight"], p[pre + "fcs.0.bias"]))
    y = F.linear(hdn, p[pre + "last_fc.weight"], p[pre + "last_fc.bias"])
    if out_act == "elu":
        return elu(y)
    if out_act == "sigmoid":
        return torch.sigmoid(y)
    return y


def dynamics_net(p, z, actions, K, det=64, return_probes=False):
    """physics_network.py:94-141 (and :156-187 for the attention probes).  z (B*K,R), actions (B*K,A) or None."""
    Kn = _n_slots(p, "dynamics_net")
    if Kn:      # PhysicsNetwork_v2_No_Sharing.forward (physics_network.py:221-232): model i on ALL rows, its rows (b, i) kept
        assert Kn == K and not return_probes
        outs = [dynamics_net(_slot_view(p, "dynamics_net", i), z, actions, K, det) for i in range(K)]
        return tuple(None if outs[0][j] is None else torch.cat([_ith(outs[i][j], K, i) for i in range(K)]) for j in range(3))
    N, R = z.shape
    B = N // K
    enc_flat = _mlp(p, "inertia_encoder", z, "elu")
    act_att = None
    if actions is not None:
        if actions.shape[-1] == 6:                                   # physics_network.py:101-102
            actions = actions[:, [0, 1, 3, 4]]
        aenc = _mlp(p, "action_encoder", actions, "elu")
        sa = torch.cat([enc_flat, aenc], -1)
        act_att = _mlp(p, "action_attention_network", sa, "sigmoid")
        enc = (_mlp(p, "action_effect_network", sa, "elu") * act_att).view(B, K, R)
    else:
        enc = enc_flat.view(B, K, R)
    pair_att = None
    if K != 1:
        idx_i = torch.tensor([i for i in range(K) for j in range(K) if j != i])
        idx_j = torch.tensor([j for i in range(K) for j in range(K) if j != i])
        pairs = torch.cat([enc[:, idx_i], enc[:, idx_j]], -1).view(B * K, K - 1, 2 * R)
        inter = _mlp(p, "pairwise_encoder_network", pairs, "elu")
        eff = _mlp(p, "interaction_effect_network", inter, "elu")
        pair_att = _mlp(p, "interaction_attention_network", inter, "sigmoid")
        total = (eff * pair_att).sum(1)
    else:
        total = torch.zeros(B, ACTION_ENC, dtype=z.dtype)
    agg = _mlp(p, "final_merge_network", torch.cat([enc.reshape(N, R), total], -1), "elu")
    deter = _mlp(p, "det_output", agg, None) if det else None
    lam1 = _mlp(p, "lambdas1_output", agg, None)
    lam2 = _mlp(p, "lambdas2_output", agg, None)
    if return_probes:
        deltas = (enc.reshape(N, R) - enc_flat).abs().sum(1, keepdim=True)
        return deter, lam1, lam2, act_att, pair_att, deltas
    return deter, lam1, lam2


# ----------------------------------------------------------------------------------------------
# Model-level steps
# ----------------------------------------------------------------------------------------------
def sample(mu, sp, eps):
    """op3_model.py:147-153."""
    return eps * softplus_std(sp) + mu


def initial_state(p, B, K, eps, det=64, sto=64):
    """op3_model.py:162-189."""
    dt = p["initial_lambdas1"].dtype
    lam1 = p["initial_lambdas1"].view(1, 1, -1).expand(B, K, sto)
    lam2 = p["initial_lambdas2"].view(1, 1, -1).expand(B, K, sto)
    return {
        "prior": {"lambdas1": torch.zeros(B, K, sto, dtype=dt),
                  "lambdas2": torch.full((B, K, sto), math.log(math.e - 1), dtype=dt)},
        "post": {"deter_state": p["inital_deter_state"].view(1, 1, -1).expand(B, K, det) if det else None,
                 "lambdas1": lam1, "lambdas2": lam2,
                 "extra_info": [torch.zeros(B * K, LSTM_SIZE, dtype=dt), torch.zeros(B * K, LSTM_SIZE, dtype=dt)],
                 "samples": sample(lam1, lam2, eps)},
    }


def refine_inputs(p, state, image, cached, beta):
    """The detached gradient inputs of the refinement network (op3_model.py:209-235) in closed form.
    ``cached`` = (colors, mask, logits, probs, P, total) of the decode of ``state``.
    Returns a (N,15,D,D), extra (N,5*sto) (the non-LN'd pieces keep their autograd history exactly as in
    the reference, the LN'd gradient pieces are detached)."""
    colors, mask, logits, probs, P, total = cached
    B, K = colors.shape[:2]
    D = colors.shape[-1]
    N = B * K
    post = state["post"]
    posterior_mask = probs / (probs.sum(1, keepdim=True) + 1e-8)
    m = mask.squeeze(2)
    leave_out = P.unsqueeze(1) - m * probs
    with torch.no_grad():
        inv = 1.0 / (P.unsqueeze(1) + 1e-12)                                    # d(-log(P+1e-12))/dP
        x_hat_grad = (m * probs * inv / B).unsqueeze(2) * (colors - image.unsqueeze(1)) / (3 * SIGMA ** 2)
        mask_grad = -(probs * inv / B).unsqueeze(2)                             # w.r.t. post-softmax mask
    g1, g2 = torch.autograd.grad(total, [post["lambdas1"], post["lambdas2"]], retain_graph=True)
    t = (N, -1, D, D)
    a = torch.cat([
        image.unsqueeze(1).expand(B, K, 3, D, D).reshape(t),
        colors.reshape(t), mask.reshape(t), logits.reshape(t), posterior_mask.reshape(t),
        layernorm2d(mask_grad.reshape(t), p["layer_norms_2d.0.gamma"], p["layer_norms_2d.0.beta"]),
        layernorm2d(P.detach().unsqueeze(1).expand(B, K, D, D).reshape(t), p["layer_norms_2d.1.gamma"], p["layer_norms_2d.1.beta"]),
        layernorm2d(leave_out.detach().reshape(t), p["layer_norms_2d.2.gamma"], p["layer_norms_2d.2.beta"]),
        layernorm2d(x_hat_grad.reshape(t), p["layer_norms_2d.3.gamma"], p["layer_norms_2d.3.beta"]),
    ], 1)
    extra = torch.cat([
        layernorm1d(g1.detach().reshape(N, -1), p["layer_norms_1d.0.scale_param"], p["layer_norms_1d.0.center_param"]),
        layernorm1d(g2.detach().reshape(N, -1), p["layer_norms_1d.1.scale_param"], p["layer_norms_1d.1.center_param"]),
        post["lambdas1"].reshape(N, -1), post["lambdas2"].reshape(N, -1), post["samples"].reshape(N, -1)], -1)
    return a, extra, dict(posterior_mask=posterior_mask, leave_out=leave_out, x_hat_grad=x_hat_grad,
                          mask_grad=mask_grad, lambdas_grad_1=g1, lambdas_grad_2=g2)


def refine(p, state, image, cached, eps, beta, D=64):
    """op3_model.py:195-264."""
    if cached is None:
        colors, mask, logits = decode(p, state, D)
        probs, P, _, _, total = get_loss(state, colors, mask, image, beta)
        cached = (colors, mask, logits, probs, P, total)
    B, K = cached[0].shape[:2]
    a, extra, _ = refine_inputs(p, state, image, cached, beta)
    h, c = state["post"]["extra_info"]
    lam1, lam2, h2, c2 = refinement_net(p, a, h.reshape(B * K, -1), c.reshape(B * K, -1), extra)
    lam1, lam2 = lam1.view(B, K, -1), lam2.view(B, K, -1)
    return {"prior": state["prior"],
            "post": {"deter_state": state["post"]["deter_state"], "lambdas1": lam1, "lambdas2": lam2,
                     "extra_info": [h2, c2], "samples": sample(lam1, lam2, eps)}}


def dynamics(p, state, actions, eps, det=64):
    """op3_model.py:269-296: new prior := new lambdas, LSTM state zeroed."""
    post = state["post"]
    B, K = post["samples"].shape[:2]
    z = post["samples"] if post["deter_state"] is None else torch.cat([post["deter_state"], post["samples"]], 2)
    acts = None if actions is None else actions.unsqueeze(1).expand(B, K, actions.shape[-1]).reshape(B * K, -1)
    deter, lam1, lam2 = dynamics_net(p, z.reshape(B * K, -1), acts, K, det)
    lam1, lam2 = lam1.view(B, K, -1), lam2.view(B, K, -1)
    h, c = post["extra_info"]
    return {"prior": {"lambdas1": lam1, "lambdas2": lam2},
            "post": {"deter_state": None if deter is None else deter.view(B, K, -1),
                     "lambdas1": lam1, "lambdas2": lam2,
                     "extra_info": [torch.zeros_like(h), torch.zeros_like(c)],
                     "samples": sample(lam1, lam2, eps)}}


def run_schedule(p, images, actions, schedule, loss_schedule, noise, K, beta=1e-2, det=64, sto=64,
                 init_state=None, D=64):
    """op3_model.py:333-434 (schedule codes 0 refine / 1 dynamics / 2 next-step refine).  ``noise`` is consumed front to back:
    one draw for the initial state (if not passed in), one per step.  Returns a dict."""
    noise = list(noise)
    B = images.shape[0] if images is not None else actions.shape[0]
    state = init_state if init_state is not None else initial_state(p, B, K, noise.pop(0), det, sto)
    colors_l, masks_l, tot, kls, clogs = [], [], [], [], []
    cur, cached = 0, None
    for i, code in enumerate(schedule):
        if code == 0:
            state = refine(p, state, images[:, cur], cached, noise.pop(0), beta, D)
        elif code == 1:
            state = dynamics(p, state, None if actions is None else actions[:, cur], noise.pop(0), det)
            cur += 1
        elif code == 2:
            # op3_model.py:368-376 next-step refinement: the action is handed to the refinement net as add_fc_input,
            # which 'size_dependent_conv' (added_fc_input_size=0) ignores (refinement_network.py:201-202)
            state = refine(p, state, images[:, cur], cached, noise.pop(0), beta, D)
            cur += 1
        else:
            raise ValueError("Invalid schedule entry: {}".format(code))
        colors, mask, logits = decode(p, state, D)
        colors_l.append(colors)
        masks_l.append(mask)
        if loss_schedule[i] != 0:
            probs, P, kl, clog, total = get_loss(state, colors, mask, images[:, cur], beta)
            tot.append(total * float(loss_schedule[i]))
            kls.append(kl * float(loss_schedule[i]))
            clogs.append(clog * float(loss_schedule[i]))
            cached = (colors, mask, logits, probs, P, total)
        else:
            cached = None
    sw = float(sum(loss_schedule))
    final_recon = (colors_l[-1] * masks_l[-1]).sum(1)
    out = {
        "colors_list": torch.stack(colors_l, 1), "masks_list": torch.stack(masks_l, 1),
        "final_recon": final_recon,
        "total_loss": sum(tot) / sw if sw else None,
        "kle_loss": sum(kls) / sw if sw else None,
        "clog_prob": sum(clogs) / sw if sw else None,
        "mse": ((final_recon - images[:, -1]) ** 2).mean() if images is not None else torch.zeros(()),
        "state": state,
    }
    return out


def train_step(p, images, actions, schedule, loss_schedule, noise, K, beta=1e-2, det=64, sto=64):
    """Forward + backward of one trainer step (op3_trainer.py:173-184).  Returns (outputs, grads dict)."""
    q = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items())
    out = run_schedule(q, images, actions, schedule, loss_schedule, noise, K, beta, det, sto)
    out["total_loss"].backward()
    grads = OrderedDict((k, (v.grad if v.grad is not None else torch.zeros_like(v))) for k, v in q.items())
    return out, grads


def clip_adam_step(params, grads, exp_avg, exp_avg_sq, step, lr=3e-4, max_norm=5.0, betas=(0.9, 0.999), eps=1e-8):
    """clip_grad_norm_(5.0) + Adam defaults (op3_trainer.py:136,188-189).  In place on params/exp_avg/exp_avg_sq
    (dicts keyed like the state_dict).  ``step`` is the 1-based step count.  Returns the pre-clip total norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
    coef = min(1.0, max_norm / (total + 1e-6))
    bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    for k in params:
        g = grads[k] * coef
        exp_avg[k].mul_(betas[0]).add_(g, alpha=1 - betas[0])
        exp_avg_sq[k].mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        denom = exp_avg_sq[k].sqrt() / math.sqrt(bc2) + eps
        params[k].addcdiv_(exp_avg[k], denom, value=-lr / bc1)
    return total


# ----------------------------------------------------------------------------------------------
# MPC cost (exps/stack_exps/mpc_stack.py:41-228 == exps/pickplace_exps/mpc_pickplace.py:50-236)
# ----------------------------------------------------------------------------------------------
def mpc_cost(goal_recon, goal_image, pred_sub_images, pred_final_recon, core_type="final_recon", aggregate="sum",
             compare_func="mse", post_process="raw", goal_latents=None, pred_latents=None):
    """Restates Cost.get_action_rankings (exps/stack_exps/mpc_stack.py:41-228).
    goal_recon (G,3,D,D), goal_image (1,3,D,D), pred_sub_images (Na,K,3,D,D), pred_final_recon (Na,3,D,D); for
    core_type 'latent' also goal_latents (G,R), pred_latents (Na,K,R).
    Returns (sorted_costs, action_index_order, goal_latent_idx) as torch tensors."""
    if core_type == "final_recon":                                  # mpc_stack.py:46-48
        goal_recon = goal_image
        pred_sub_images = pred_final_recon.unsqueeze(1)
    costs = []
    for g in range(goal_recon.shape[0]):
        if core_type == "latent":                                   # compare_latents :73-77, mse over the last axis :103-107
            d = ((goal_latents[g].view(1, 1, -1) - pred_latents) ** 2).mean(-1)
        elif compare_func == "mse":                                 # image_mse :109-127
            d = ((goal_recon[g].view(1, 1, *goal_recon.shape[1:]) - pred_sub_images) ** 2).mean((-1, -2, -3))  # (Na,K)
        elif compare_func == "psuedo_intersect":                    # :84-92
            gr = goal_recon[g]
            m1, m2 = (gr > 0.01).float(), (pred_sub_images > 0.01).float()
            union = m1.sum((-3, -2, -1)) + m2.sum((-3, -2, -1))
            ti = (m1 * m2 * ((gr - pred_sub_images) ** 2 < 0.08).float()).sum((-3, -2, -1))
            d = 1 - ti / union
        else:
            raise KeyError(compare_func)
        if post_process == "negative_exp":                          # :94-100
            d = -torch.exp(-d)
        elif post_process != "raw":
            raise KeyError(post_process)
        costs.append(d.min(-1)[0])
    costs = torch.stack(costs)                                      # (G,Na)
    if aggregate == "sum":
        sorted_costs, order = costs.sum(0).sort()
        return sorted_costs, order, torch.zeros(len(order), dtype=torch.long)
    mins, gidx = costs.min(0)
    sorted_costs, order = mins.sort()
    return sorted_costs, order, gidx
