"""Standalone (inference-only) forwards of the three sub-networks through the C ABI, used by
DecoderNetwork.forward, PhysicsNetwork.forward / attention probes and OP3Model.decode /
get_all_activation_values.  Training never comes through here (see engine.py)."""
import ctypes as C

import torch

from op3_b200 import _lib
from op3_b200._lib import Op3Dims, Op3Params, check, ptr


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("op3_b200: %s is on %s; the OP3 kernels are CUDA-only (no CPU fallback)" % (what, t.device))


def _on_tensor_device(fn):
    """Run `fn(net, x, ...)` with x's device current: launches and attribute caches are per device."""
    import functools

    @functools.wraps(fn)
    def wrapper(net, x, *args, **kw):
        _require_cuda(x, "input")
        with torch.cuda.device(x.device):
            return fn(net, x, *args, **kw)
    return wrapper


@_on_tensor_device
def decoder_forward(net, latents, engine=None):
    """DecoderNetwork.forward (decoder_network.py:83-88): latents (N,R) -> (mask_logits (N,1,D,D), colors (N,3,D,D))."""
    lib = _lib.load()
    _require_cuda(latents, "latents")
    z = latents.detach().contiguous().float()
    N, R = z.shape
    D = net.input_width
    convs = net.broadcast_net.conv_layers
    if convs[0].in_channels != R + 2:
        raise ValueError("latent size %d does not match the decoder (expects %d)" % (R, convs[0].in_channels - 2))
    st = _stream(z.device)
    d = Op3Dims(B=N, K=1, D=D, Rd=R - 64, Rs=64, A=0, beta=0.0)
    if engine is not None:
        prep = engine.prepared(st)
    else:
        ps = Op3Params()
        keep = []
        for i, c in enumerate(convs):
            w, b = c.weight.detach().contiguous(), c.bias.detach().contiguous()
            _require_cuda(w, "decoder weights")
            keep += [w, b]
            ps.dec_conv_w[i], ps.dec_conv_b[i] = w.data_ptr(), b.data_ptr()
        prep = torch.empty(lib.op3_prepared_floats(C.byref(d)), device=z.device)
        check(lib.op3_prepare(C.byref(d), C.byref(ps), ptr(prep), st), "op3_prepare")
    acts = torch.empty(lib.op3_decoder_acts_floats(C.byref(d), N), device=z.device)
    out = torch.empty(N, D, D, 4, device=z.device)
    check(lib.op3_decoder_fwd(C.byref(d), N, ptr(prep), ptr(z), ptr(acts), ptr(out), st), "op3_decoder_fwd")
    nchw = out.permute(0, 3, 1, 2)
    return nchw[:, 3:4].contiguous(), nchw[:, :3].contiguous()


def _mlp_ptrs(struct_mlp, mlp, keep):
    for lin, field in [(mlp.fcs[0], struct_mlp.fc0), (mlp.last_fc, struct_mlp.last)]:
        w, b = lin.weight.detach().contiguous(), lin.bias.detach().contiguous()
        _require_cuda(w, "dynamics weights")
        keep += [w, b]
        field.w, field.b = w.data_ptr(), b.data_ptr()


@_on_tensor_device
def physics_forward(net, sampled_state, actions, K=None, engine=None):
    """PhysicsNetwork.forward (physics_network.py:94-141) + attention probes (:156-187).
    Returns (deter or None, lambdas1, lambdas2, act_att (N,1) or None, pair_att (N,K-1,1) or None, deltas (N,1))."""
    lib = _lib.load()
    _require_cuda(sampled_state, "sampled_state")
    K = net.K if K is None else K
    z = sampled_state.detach().contiguous().float()
    N, R = z.shape
    if N % K:
        raise ValueError("rows (%d) must be a multiple of K (%d)" % (N, K))
    B = N // K
    dev = z.device
    st = _stream(dev)
    A = net.action_size if actions is not None else 0
    d = Op3Dims(B=B, K=K, D=64, Rd=net.det_size, Rs=net.sto_size, A=A, beta=0.0)
    acts_b = None
    if actions is not None:
        a = actions.detach().float()
        if net.action_size == 4 and a.shape[-1] == 6:
            a = a[:, [0, 1, 3, 4]]
        acts_b = a.view(B, K, -1)[:, 0].contiguous()      # rows of one sample share the action (op3_model.py:272)
    keep = []
    if engine is not None:
        ps = engine.params_struct()
    else:
        ps = Op3Params()
        names = [("inertia_encoder", "dyn_inertia"), ("pairwise_encoder_network", "dyn_pairwise"),
                 ("interaction_effect_network", "dyn_inter_effect"), ("interaction_attention_network", "dyn_inter_att"),
                 ("final_merge_network", "dyn_merge"), ("lambdas1_output", "dyn_lam1_out"), ("lambdas2_output", "dyn_lam2_out")]
        if net.action_size > 0:
            names += [("action_encoder", "dyn_action_enc"), ("action_effect_network", "dyn_action_effect"),
                      ("action_attention_network", "dyn_action_att")]
        if net.det_size:
            names.append(("det_output", "dyn_det_out"))
        for attr, field in names:
            _mlp_ptrs(getattr(ps, field), getattr(net, attr), keep)
    acts = torch.empty(lib.op3_dynamics_acts_floats(C.byref(d)), device=dev)
    out_z = torch.empty(N, R, device=dev)
    l1, l2 = torch.empty(N, net.sto_size, device=dev), torch.empty(N, net.sto_size, device=dev)
    check(lib.op3_dynamics_fwd(C.byref(d), C.byref(ps), ptr(z), ptr(acts_b), ptr(acts), ptr(out_z) if net.det_size else None,
                               ptr(l1), ptr(l2), st), "op3_dynamics_fwd")
    act_att = torch.empty(N, 1, device=dev) if A > 0 else None
    pair_att = torch.empty(N, K - 1, 1, device=dev) if K > 1 else None
    deltas = torch.empty(N, 1, device=dev)
    check(lib.op3_dynamics_probes(C.byref(d), ptr(acts), ptr(act_att), ptr(pair_att), ptr(deltas), st), "op3_dynamics_probes")
    deter = out_z[:, :net.det_size].contiguous() if net.det_size else None
    return deter, l1, l2, act_att, pair_att, deltas


# reference channel (op3_model.py:219-229 order + coords 15,16) stored at each of the 20 packed channels; -1 = zero
_REF_CHANNEL = [3, 4, 5, 7, 6, 8, 9, 11, 12, 13, 14, -1, 0, 1, 2, 10, 15, 16, -1, -1]


@_on_tensor_device
def refinement_forward(net, a, h, c, extra):
    """RefinementNetwork.forward (refinement_network.py:189-219): a (N,15,D,D), h/c (N,128) or (1,N,128), extra (N,5*Rs)
    -> (lambdas1 (N,Rs), lambdas2 (N,Rs), h (1,N,128), c (1,N,128))."""
    lib = _lib.load()
    _require_cuda(a, "input")
    dev = a.device
    N, _, D, _ = a.shape
    Rs = net.output_size
    st = _stream(dev)
    x = torch.cat([a.detach().float(), net.coords.to(dev).float().expand(N, 2, D, D)], 1)          # (N,17,D,D)
    idx = torch.tensor([max(ch, 0) for ch in _REF_CHANNEL], device=dev)
    keep = torch.tensor([1.0 if ch >= 0 else 0.0 for ch in _REF_CHANNEL], device=dev).view(1, 20, 1, 1)
    packed = (x[:, idx] * keep).view(N, 5, 4, D, D).permute(0, 1, 3, 4, 2).contiguous()       # [N][5][D][D][4]
    ps = Op3Params()
    keepalive = []

    def setp(field, lin):
        w, b = lin.weight.detach().contiguous(), lin.bias.detach().contiguous()
        _require_cuda(w, "refinement weights")
        keepalive.extend([w, b])
        field.w, field.b = w.data_ptr(), b.data_ptr()
    for i, cv in enumerate(net.conv_layers):
        w, b = cv.weight.detach().contiguous(), cv.bias.detach().contiguous()
        _require_cuda(w, "refinement weights")
        keepalive.extend([w, b])
        ps.ref_conv_w[i], ps.ref_conv_b[i] = w.data_ptr(), b.data_ptr()
    setp(ps.ref_fc, net.fc_layers[0])
    setp(ps.ref_last_fc, net.last_fc)
    setp(ps.ref_last_fc2, net.last_fc2)
    for name, field in [("weight_ih_l0", "lstm_w_ih"), ("weight_hh_l0", "lstm_w_hh"), ("bias_ih_l0", "lstm_b_ih"),
                        ("bias_hh_l0", "lstm_b_hh")]:
        t = getattr(net.lstm, name).detach().contiguous()
        keepalive.append(t)
        setattr(ps, field, t.data_ptr())
    d = Op3Dims(B=N, K=1, D=D, Rd=0, Rs=Rs, A=0, beta=0.0)
    prep = torch.empty(lib.op3_prepared_floats(C.byref(d)), device=dev)
    check(lib.op3_prepare(C.byref(d), C.byref(ps), ptr(prep), st), "op3_prepare")
    acts = torch.empty(lib.op3_refine_acts_floats(C.byref(d)), device=dev)
    hh = h.detach().reshape(N, -1).contiguous().float()
    cc = c.detach().reshape(N, -1).contiguous().float()
    ex = extra.detach().contiguous().float()
    l1, l2 = torch.empty(N, Rs, device=dev), torch.empty(N, Rs, device=dev)
    h2, c2 = torch.empty(N, hh.shape[1], device=dev), torch.empty(N, hh.shape[1], device=dev)
    check(lib.op3_refine_net_fwd(C.byref(d), C.byref(ps), ptr(prep), ptr(packed), ptr(ex), ptr(hh), ptr(cc), ptr(acts),
                                 ptr(l1), ptr(l2), ptr(h2), ptr(c2), st), "op3_refine_net_fwd")
    return l1, l2, h2.unsqueeze(0), c2.unsqueeze(0)
