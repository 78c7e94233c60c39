"""Schedule executor: runs OP3Model.run_schedule (reference op3/torch/op3_modules/op3_model.py:333-434) and its
backward entirely through the C-ABI kernels of libop3_b200.so.

PyTorch is used for device memory (torch.empty), streams and the autograd hand-over only.  The unrolled schedule is
recorded on a small tape (states, decodes, steps); backward walks it in reverse, calling the *_bwd entry points.
Because refine() detaches every gradient input it computes (op3_model.py:225-234) no double-backward is needed:
the refinement "gradient inputs" are first-order closed forms produced in the forward pass.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import Op3Dims, Op3Params, Op3MixtureIO, Op3RefineIO, check, ptr

LSTM = 128
PRIOR_SP = math.log(math.e - 1.0)   # softplus^-1(1): initial prior lambdas2 (op3_model.py:179)


def _pad4(n):
    return (n + 3) & ~3


# state_dict key -> path in the Op3Params struct
def _param_field_map(action_dim, det):
    m = {"initial_lambdas1": ("initial_lambdas1",), "initial_lambdas2": ("initial_lambdas2",)}
    if det:
        m["inital_deter_state"] = ("inital_deter_state",)
    for i in range(3):
        m["refinement_net.conv_layers.%d.weight" % i] = ("ref_conv_w", i)
        m["refinement_net.conv_layers.%d.bias" % i] = ("ref_conv_b", i)
    m["refinement_net.fc_layers.0.weight"] = ("ref_fc", "w")
    m["refinement_net.fc_layers.0.bias"] = ("ref_fc", "b")
    m["refinement_net.lstm.weight_ih_l0"] = ("lstm_w_ih",)
    m["refinement_net.lstm.weight_hh_l0"] = ("lstm_w_hh",)
    m["refinement_net.lstm.bias_ih_l0"] = ("lstm_b_ih",)
    m["refinement_net.lstm.bias_hh_l0"] = ("lstm_b_hh",)
    for k, f in [("last_fc", "ref_last_fc"), ("last_fc2", "ref_last_fc2")]:
        m["refinement_net.%s.weight" % k] = (f, "w")
        m["refinement_net.%s.bias" % k] = (f, "b")
    mlps = [("inertia_encoder", "dyn_inertia"), ("pairwise_encoder_network", "dyn_pairwise"),
            ("interaction_effect_network", "dyn_inter_effect"), ("interaction_attention_network", "dyn_inter_att"),
            ("final_merge_network", "dyn_merge"), ("lambdas1_output", "dyn_lam1_out"), ("lambdas2_output", "dyn_lam2_out")]
    if action_dim > 0:
        mlps += [("action_encoder", "dyn_action_enc"), ("action_effect_network", "dyn_action_effect"),
                 ("action_attention_network", "dyn_action_att")]
    if det:
        mlps.append(("det_output", "dyn_det_out"))
    for k, f in mlps:
        m["dynamics_net.%s.fcs.0.weight" % k] = (f, "fc0", "w")
        m["dynamics_net.%s.fcs.0.bias" % k] = (f, "fc0", "b")
        m["dynamics_net.%s.last_fc.weight" % k] = (f, "last", "w")
        m["dynamics_net.%s.last_fc.bias" % k] = (f, "last", "b")
    for i in range(5):
        m["decode_net.broadcast_net.conv_layers.%d.weight" % i] = ("dec_conv_w", i)
        m["decode_net.broadcast_net.conv_layers.%d.bias" % i] = ("dec_conv_b", i)
    for i in range(4):
        m["layer_norms_2d.%d.gamma" % i] = ("ln2d_gamma", i)
        m["layer_norms_2d.%d.beta" % i] = ("ln2d_beta", i)
    for i in range(2):
        m["layer_norms_1d.%d.scale_param" % i] = ("ln1d_scale", i)
        m["layer_norms_1d.%d.center_param" % i] = ("ln1d_center", i)
    return m


def _split_slot(name):
    """'decode_net.models.2.broadcast_net...' -> (2, 'decode_net.broadcast_net...'); names of shared tensors -> (None, name).
    The *_No_Sharing networks keep one sub-network per slot under `models.<k>` (decoder_network.py:113-131)."""
    parts = name.split(".")
    if len(parts) > 2 and parts[1] == "models" and parts[2].isdigit():
        return int(parts[2]), ".".join(parts[:1] + parts[3:])
    return None, name


def _set_field(struct, path, value):
    obj = struct
    for p in path[:-1]:
        obj = getattr(obj, p)
    last = path[-1]
    if isinstance(last, int):
        obj[last] = value
    else:
        setattr(obj, last, value)


class _State:
    """Latent state of all N = B*K slots (reference hidden_state dict, op3_model.py:176-188)."""
    __slots__ = ("z", "l1", "l2", "eps", "h", "c", "prior_kind", "prior_ref", "p1", "p2",
                 "dz", "dl1", "dl2", "dh", "dc")

    def __init__(self):
        self.dz = self.dl1 = self.dl2 = self.dh = self.dc = None
        self.prior_ref = None

    def prior_tensors(self):
        if self.prior_kind == "alias":
            return self.l1, self.l2
        if self.prior_kind == "node":
            return self.prior_ref.l1, self.prior_ref.l2
        return self.p1, self.p2

    def prior_owner(self):
        """State whose lambda gradients receive d/d prior (None: the prior is a constant)."""
        if self.prior_kind == "alias":
            return self
        if self.prior_kind == "node":
            return self.prior_ref
        return None


class _Decode:
    __slots__ = ("state", "acts", "dec_out", "io", "keep", "weight", "t_index", "has_stats", "refine_after", "losses", "zp")


class _Step:
    __slots__ = ("kind", "src", "dst", "dec", "acts", "dz", "io", "keep", "actions")


class Tape:
    pass


class Engine:
    def __init__(self, model):
        self.model = model
        self.lib = _lib.load()
        self._flat = None
        self._offsets = None
        self._names = None
        self._pstruct = None
        self._pstruct_key = None
        self._prep = None
        self._prep_key = None
        self.manual_version = 0         # bumped by the fused optimizer (writes through raw pointers)
        self._mix_sync = {}             # (device, stream) -> zeroed device word of the single-launch mixture forward
        self._w_cache = {}              # (loss schedule, device) -> normalised loss weights on the device
        self.last_flat_grad = None
        # *_No_Sharing networks: one weight set per slot, same kernels launched once per slot
        self.ns_dec = bool(getattr(model.decode_net, "no_sharing", False))
        self.ns_ref = bool(getattr(model.refinement_net, "no_sharing", False))
        self.ns_dyn = bool(getattr(model.dynamics_net, "no_sharing", False))
        self._pstructs = None
        self._preps = None

    @property
    def n_slots(self):
        """Number of distinct weight sets: K when any sub-network is a *_No_Sharing one, else 1."""
        return self.model.K if (self.ns_dec or self.ns_ref or self.ns_dyn) else 1

    # ------------------------------------------------------------------ parameters
    def _named_params(self):
        return list(self.model.named_parameters())

    def flat_params(self):
        """All parameters as views of ONE flat fp32 CUDA buffer (fused Adam / NCCL allreduce operate on it)."""
        named = self._named_params()
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("op3_b200: model parameters are on %s; the OP3 kernels are CUDA-only (no CPU fallback). "
                               "Move the model with .cuda() / ptu.set_gpu_mode(True)." % dev)
        ok = self._flat is not None and self._flat.device == dev and self._names == [n for n, _ in named]
        if ok:
            base = self._flat.data_ptr()
            for (n, p), o in zip(named, self._offsets):
                if p.data_ptr() != base + 4 * o or p.dtype != torch.float32:
                    ok = False
                    break
        if not ok:
            offs, o = [], 0
            for _, p in named:
                offs.append(o)
                o += _pad4(p.numel())
            flat = torch.zeros(o, device=dev, dtype=torch.float32)
            for (n, p), off in zip(named, offs):
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
            self._flat, self._offsets, self._names = flat, offs, [n for n, _ in named]
            self._pstruct_key = None
            self._prep_key = None
        return self._flat

    def _structs_for(self, flat):
        """One Op3Params struct per weight set, pointing into `flat` (parameter buffer or a gradient buffer of the same
        layout).  Struct k holds the shared tensors plus the `models.<k>` tensors of the *_No_Sharing networks."""
        m = self.model
        fmap = _param_field_map(m.action_dim, m.det_size)
        structs = [Op3Params() for _ in range(self.n_slots)]
        base = flat.data_ptr()
        for name, off in zip(self._names, self._offsets):
            slot, key = _split_slot(name)
            if key not in fmap or (slot is not None and slot >= len(structs)):
                raise RuntimeError("op3_b200: unexpected parameter %r" % name)
            for k, s in enumerate(structs):
                if slot is None or slot == k:
                    _set_field(s, fmap[key], base + 4 * off)
        return structs

    def _struct_for(self, flat):
        """Op3Params struct of weight set 0 (the only one for the shared-weight networks)."""
        return self._structs_for(flat)[0]

    def mix_sync(self, dev, stream):
        """The self-resetting 'samples done' word of op3_mixture_fwd's single-launch path (Op3MixtureIO.sync)."""
        key = (dev.index, stream)
        t = self._mix_sync.get(key)
        if t is None:
            t = self._mix_sync[key] = torch.zeros(4, device=dev, dtype=torch.int32)
        return t

    def _loss_weights(self, lw, sw, dev):
        """loss_schedule / sum(loss_schedule) as a device tensor, cached per schedule (an H2D copy per call would also
        be illegal inside a CUDA-graph capture)."""
        key = (tuple(lw), dev.index)
        w = self._w_cache.get(key)
        if w is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("op3_b200: run this schedule once eagerly before capturing it in a CUDA graph")
            w = self._w_cache[key] = torch.tensor(lw, device=dev, dtype=torch.float32) / sw
        return w

    def dims(self, B):
        m = self.model
        return Op3Dims(B=B, K=m.K, D=m.image_size, Rd=m.det_size, Rs=m.sto_size, A=m.action_dim, beta=float(m.beta))

    def params_structs(self):
        flat = self.flat_params()
        key = (flat.data_ptr(), self.n_slots)
        if self._pstruct_key != key:
            self._pstructs = self._structs_for(flat)
            self._pstruct = self._pstructs[0]
            self._pstruct_key = key
        return self._pstructs

    def params_struct(self):
        return self.params_structs()[0]

    def _version_key(self):
        return (self._flat.data_ptr(), sum(p._version for _, p in self._named_params()), self.manual_version,
                self.lib.op3_get_precision())

    def prepared_all(self, stream):
        """Derived weights (op3_prepare) of every weight set, refreshed whenever the parameters changed."""
        pss = self.params_structs()
        nprep = len(pss) if (self.ns_dec or self.ns_ref) else 1      # only the conv networks have derived weights
        key = self._version_key() + (nprep,)
        if self._preps is None or self._prep_key != key or self._preps[0].device != self._flat.device:
            d = self.dims(1)
            n = self.lib.op3_prepared_floats(C.byref(d))
            if n == 0:
                check(-1, "op3_prepared_floats")
            if (self._preps is None or len(self._preps) != nprep or self._preps[0].numel() != n
                    or self._preps[0].device != self._flat.device):
                self._preps = [torch.empty(n, device=self._flat.device, dtype=torch.float32) for _ in range(nprep)]
            for k in range(nprep):
                check(self.lib.op3_prepare(C.byref(d), C.byref(pss[k]), ptr(self._preps[k]), stream), "op3_prepare")
            self._prep = self._preps[0]
            self._prep_key = key
        return self._preps

    def prepared(self, stream):
        return self.prepared_all(stream)[0]

    # ------------------------------------------------------------------ forward
    def forward(self, images, actions, init_state, schedule, loss_schedule, noise_fn, need_grad):
        """Returns (outputs dict, tape or None).  images (B,T,3,D,D) float32 cuda or None; actions (B,T,A) or None."""
        ref = images if images is not None else actions if actions is not None else (
            init_state["post"]["samples"] if init_state is not None else None)
        if ref is None:
            raise ValueError("Need either images or actions")
        if ref.device.type != "cuda":
            raise RuntimeError("op3_b200: inputs are on %s; the OP3 kernels are CUDA-only (no CPU fallback)" % ref.device)
        pdev = self.flat_params().device
        if ref.device != pdev:
            raise RuntimeError("op3_b200: inputs are on %s but the model parameters are on %s (one process per GPU; "
                               "nn.DataParallel replication is not supported)" % (ref.device, pdev))
        with torch.cuda.device(ref.device):       # launches, attribute caches and streams belong to the tensors' device
            return self._forward(images, actions, init_state, schedule, loss_schedule, noise_fn, need_grad)

    def _forward(self, images, actions, init_state, schedule, loss_schedule, noise_fn, need_grad):
        m, lib = self.model, self.lib
        if images is not None:
            B, dev = images.shape[0], images.device
        elif actions is not None:
            B, dev = actions.shape[0], actions.device
        elif init_state is not None:
            B, dev = init_state["post"]["samples"].shape[0], init_state["post"]["samples"].device
        else:
            raise ValueError("Need either images or actions")
        if dev.type != "cuda":
            raise RuntimeError("op3_b200: inputs are on %s; the OP3 kernels are CUDA-only (no CPU fallback)" % dev)
        K, D, Rd, Rs = m.K, m.image_size, m.det_size, m.sto_size
        R, N = Rd + Rs, B * K
        schedule = [int(s) for s in schedule]
        lw = [float(w) for w in loss_schedule]
        T1 = len(schedule)
        for s in schedule:
            if s not in (0, 1, 2):
                raise ValueError("Invalid schedule entry: {}".format(s))
        st = torch.cuda.current_stream(dev).cuda_stream
        d = self.dims(B)
        pss = self.params_structs()
        preps = self.prepared_all(st)
        ps, prep = pss[0], preps[0]
        f32 = dict(device=dev, dtype=torch.float32)
        DD = D * D
        # *_No_Sharing networks: slot k's rows (b, k) are gathered slot-major, the kernels run once per slot on B rows with
        # weight set k, and the K results stay slot-major (row k*B + b) -- the reference's torch.cat row order, which the rest
        # of the model reads as (b, k)-major again (decoder_network.py:146-162; kept as is)
        ns_dec, ns_ref, ns_dyn = self.ns_dec, self.ns_ref, self.ns_dyn
        d1 = Op3Dims(B=B, K=1, D=D, Rd=Rd, Rs=Rs, A=m.action_dim, beta=float(m.beta))     # one slot's B rows

        def slot_major(t):
            return t.view(B, K, *t.shape[1:]).transpose(0, 1).contiguous()                # (K, B, ...)

        if images is not None:
            images = images.contiguous().float()
            if images.shape[-1] != D or images.shape[-2] != D:
                raise ValueError("images must be (B,T,3,%d,%d), got %s" % (D, D, tuple(images.shape)))
            img_bstride = images.stride(0)
        acts_t = None
        if actions is not None:
            a = actions.float()
            if m.action_dim == 4 and a.shape[-1] == 6:            # physics_network.py:101-102 (columns 0,1,3,4)
                a = torch.cat([a[..., 0:2], a[..., 3:5]], -1)      # slices, not an index tensor: CUDA-graph capturable
            if m.action_dim > 0:
                if a.shape[-1] != m.action_dim:
                    raise ValueError("actions last dim %d != action_dim %d" % (a.shape[-1], m.action_dim))
                acts_t = a.transpose(0, 1).contiguous()            # (T,B,A): one contiguous [B][A] block per step
        tape = Tape() if need_grad else None

        # ---- initial state (op3_model.py:162-189) or the one passed in
        s0 = _State()
        if init_state is None:
            s0.l1 = m.initial_lambdas1.detach().view(1, Rs).expand(N, Rs).contiguous()
            s0.l2 = m.initial_lambdas2.detach().view(1, Rs).expand(N, Rs).contiguous()
            s0.z = torch.empty(N, R, **f32)
            if Rd:
                s0.z[:, :Rd] = m.inital_deter_state.detach().view(1, Rd)
            s0.eps = noise_fn(B, K, Rs, dev).reshape(N, Rs).contiguous()
            check(lib.op3_sample_fwd(C.byref(d), N, ptr(s0.l1), ptr(s0.l2), ptr(s0.eps), None, ptr(s0.z), st), "sample")
            s0.h = torch.zeros(N, LSTM, **f32)
            s0.c = torch.zeros(N, LSTM, **f32)
            s0.prior_kind = "const"
            s0.p1 = torch.zeros(N, Rs, **f32)
            s0.p2 = torch.full((N, Rs), PRIOR_SP, **f32)
            from_params = True
        else:
            post, prior = init_state["post"], init_state["prior"]
            if post["samples"].shape[0] != B:
                raise ValueError("initial_hidden_state batch %d != input batch %d" % (post["samples"].shape[0], B))
            s0.l1 = post["lambdas1"].detach().reshape(N, Rs).contiguous().float()
            s0.l2 = post["lambdas2"].detach().reshape(N, Rs).contiguous().float()
            s0.z = torch.empty(N, R, **f32)
            if Rd:
                s0.z[:, :Rd] = post["deter_state"].detach().reshape(N, Rd)
            s0.z[:, Rd:] = post["samples"].detach().reshape(N, Rs)
            s0.eps = torch.zeros(N, Rs, **f32)      # unknown noise of a passed-in sample; only matters for gradients
            s0.h = post["extra_info"][0].detach().reshape(N, LSTM).contiguous().float()
            s0.c = post["extra_info"][1].detach().reshape(N, LSTM).contiguous().float()
            s0.prior_kind = "tensor"
            s0.p1 = prior["lambdas1"].detach().reshape(N, Rs).contiguous().float()
            s0.p2 = prior["lambdas2"].detach().reshape(N, Rs).contiguous().float()
            from_params = False

        colors_list = torch.empty(T1, B, K, 3, D, D, **f32)
        masks_list = torch.empty(T1, B, K, 1, D, D, **f32)
        final_recon = torch.empty(B, 3, D, D, **f32)
        loss_rows = torch.zeros(max(T1, 1), 4, **f32)

        mix_sync = self.mix_sync(dev, st)
        n_dec_acts = lib.op3_decoder_acts_floats(C.byref(d1), B) if ns_dec else lib.op3_decoder_acts_floats(C.byref(d), N)
        n_stats = lib.op3_mixture_stats_floats(C.byref(d))
        n_ref_acts = lib.op3_refine_acts_floats(C.byref(d1 if ns_ref else d))
        n_dyn_acts = lib.op3_dynamics_acts_floats(C.byref(d))
        dec_scratch = torch.empty(lib.op3_decoder_scratch_floats(C.byref(d), N), **f32)

        def decode(state, t_index, image, weight, refine_after, is_last):
            rec = _Decode()
            rec.state, rec.weight, rec.t_index, rec.refine_after = state, weight, t_index, refine_after
            rec.dec_out = torch.empty(N, DD, 4, **f32)
            if not ns_dec:
                rec.acts = torch.empty(n_dec_acts, **f32)
                check(lib.op3_decoder_fwd(C.byref(d), N, ptr(prep), ptr(state.z), ptr(rec.acts), ptr(rec.dec_out), st), "decoder_fwd")
            else:
                rec.zp = slot_major(state.z)
                rec.acts = [torch.empty(n_dec_acts, **f32) for _ in range(K)]
                out_v = rec.dec_out.view(K, B, DD, 4)
                for k in range(K):
                    check(lib.op3_decoder_fwd(C.byref(d1), B, ptr(preps[k]), ptr(rec.zp[k]), ptr(rec.acts[k]), ptr(out_v[k]), st),
                          "decoder_fwd")
            want_stats = (weight != 0.0) or refine_after or (is_last and images is not None)
            rec.has_stats = want_stats
            io = Op3MixtureIO()
            keep = []
            io.dec_out = ptr(rec.dec_out)
            if image is not None:
                io.image, io.image_bstride = image.data_ptr(), img_bstride
                keep.append(image)
            if is_last and images is not None:
                io.mse_image, io.mse_image_bstride = images[:, -1].data_ptr(), img_bstride
            p1, p2 = state.prior_tensors()
            io.post_l1, io.post_l2, io.prior_l1, io.prior_l2 = ptr(state.l1), ptr(state.l2), ptr(p1), ptr(p2)
            if t_index is not None:
                io.mask = masks_list[t_index].data_ptr()
                io.colors = colors_list[t_index].data_ptr()
            if is_last:
                io.final_recon = final_recon.data_ptr()
            if want_stats:
                if image is None:
                    raise ValueError("a loss / refinement step needs images")
                stats = torch.empty(n_stats, **f32)
                rec.losses = loss_rows[t_index] if t_index is not None else torch.empty(4, **f32)
                io.stats, io.losses = ptr(stats), rec.losses.data_ptr()
                io.sync = mix_sync.data_ptr()
                keep += [stats, rec.losses]
            if refine_after:
                mix1 = torch.empty(N, DD, 4, **f32)
                mix2 = torch.empty(N, DD, 4, **f32)
                smp = torch.empty(B, DD, 4, **f32)
                seed = torch.empty(N, DD, 4, **f32)
                io.mix1, io.mix2, io.smp, io.seed = ptr(mix1), ptr(mix2), ptr(smp), ptr(seed)
                keep += [mix1, mix2, smp, seed]
            check(lib.op3_mixture_fwd(C.byref(d), C.byref(ps), C.byref(io), 1 if want_stats else 0, st), "mixture_fwd")
            rec.io, rec.keep = io, keep
            return rec

        def refine(state, dec, eps_new):
            step = _Step()
            step.kind, step.src, step.dec = "refine", state, dec
            step.dz = torch.empty(N, R, **f32)
            seed = dec.keep[-1]
            if not ns_dec:
                check(lib.op3_decoder_bwd(C.byref(d), N, ptr(prep), ptr(state.z), ptr(dec.acts), ptr(seed), ptr(dec_scratch),
                                          ptr(step.dz), 0, None, st), "decoder_bwd(latent)")
            else:
                seed_v, dzp = seed.view(K, B, DD, 4), torch.empty(K, B, R, **f32)
                for k in range(K):
                    check(lib.op3_decoder_bwd(C.byref(d1), B, ptr(preps[k]), ptr(dec.zp[k]), ptr(dec.acts[k]), ptr(seed_v[k]),
                                              ptr(dec_scratch), ptr(dzp[k]), 0, None, st), "decoder_bwd(latent)")
                step.dz = dzp.transpose(0, 1).reshape(N, R)
            new = _State()
            new.l1, new.l2 = torch.empty(N, Rs, **f32), torch.empty(N, Rs, **f32)
            new.h, new.c = torch.empty(N, LSTM, **f32), torch.empty(N, LSTM, **f32)
            new.z, new.eps = torch.empty(N, R, **f32), eps_new
            p1, p2 = state.prior_tensors()
            if not ns_ref:
                step.acts = torch.empty(n_ref_acts, **f32)
                io = Op3RefineIO()
                io.dec_out, io.mix1, io.mix2, io.smp = dec.io.dec_out, dec.io.mix1, dec.io.mix2, dec.io.smp
                io.dz, io.z, io.l1, io.l2, io.eps = ptr(step.dz), ptr(state.z), ptr(state.l1), ptr(state.l2), ptr(state.eps)
                io.prior1, io.prior2, io.prior_is_post = ptr(p1), ptr(p2), 1 if state.prior_kind == "alias" else 0
                io.h, io.c, io.acts = ptr(state.h), ptr(state.c), ptr(step.acts)
                io.l1_out, io.l2_out, io.h_out, io.c_out = ptr(new.l1), ptr(new.l2), ptr(new.h), ptr(new.c)
                check(lib.op3_refine_fwd(C.byref(d), C.byref(ps), ptr(prep), C.byref(io), st), "refine_fwd")
            else:
                # refinement_network.py:291-307: network k sees the rows (b, k) of every input; outputs stay slot-major
                mix1, mix2 = dec.keep[-4], dec.keep[-3]
                pin = [slot_major(t) for t in (dec.dec_out, mix1, mix2, step.dz, state.z, state.l1, state.l2, state.eps,
                                               p1, p2, state.h, state.c)]
                step.acts = [torch.empty(n_ref_acts, **f32) for _ in range(K)]
                outs = [t.view(K, B, -1) for t in (new.l1, new.l2, new.h, new.c)]
                io = []
                for k in range(K):
                    o = Op3RefineIO()
                    o.dec_out, o.mix1, o.mix2, o.smp = ptr(pin[0][k]), ptr(pin[1][k]), ptr(pin[2][k]), dec.io.smp
                    o.dz, o.z, o.l1, o.l2, o.eps = ptr(pin[3][k]), ptr(pin[4][k]), ptr(pin[5][k]), ptr(pin[6][k]), ptr(pin[7][k])
                    o.prior1, o.prior2, o.prior_is_post = ptr(pin[8][k]), ptr(pin[9][k]), 1 if state.prior_kind == "alias" else 0
                    o.h, o.c, o.acts = ptr(pin[10][k]), ptr(pin[11][k]), ptr(step.acts[k])
                    o.l1_out, o.l2_out, o.h_out, o.c_out = ptr(outs[0][k]), ptr(outs[1][k]), ptr(outs[2][k]), ptr(outs[3][k])
                    check(lib.op3_refine_fwd(C.byref(d1), C.byref(pss[k]), ptr(preps[k]), C.byref(o), st), "refine_fwd")
                    io.append(o)
                step.keep = pin
            check(lib.op3_sample_fwd(C.byref(d), N, ptr(new.l1), ptr(new.l2), ptr(new.eps), ptr(state.z), ptr(new.z), st), "sample")
            # the prior is carried over unchanged (op3_model.py:255)
            if state.prior_kind == "alias":
                new.prior_kind, new.prior_ref = "node", state
            elif state.prior_kind == "node":
                new.prior_kind, new.prior_ref = "node", state.prior_ref
            else:
                new.prior_kind, new.p1, new.p2 = state.prior_kind, state.p1, state.p2
            step.io, step.dst = io, new
            return step

        def dynamics(state, act, eps_new):
            step = _Step()
            step.kind, step.src, step.actions = "dynamics", state, act
            new = _State()
            new.l1, new.l2 = torch.empty(N, Rs, **f32), torch.empty(N, Rs, **f32)
            new.z, new.eps = torch.empty(N, R, **f32), eps_new
            new.h, new.c = torch.zeros(N, LSTM, **f32), torch.zeros(N, LSTM, **f32)   # op3_model.py:277-278
            if not ns_dyn:
                step.acts = torch.empty(n_dyn_acts, **f32)
                check(lib.op3_dynamics_fwd(C.byref(d), C.byref(ps), ptr(state.z), ptr(act), ptr(step.acts),
                                           ptr(new.z) if Rd else None, ptr(new.l1), ptr(new.l2), st), "dynamics_fwd")
            else:
                # physics_network.py:221-232: model i runs on ALL rows; its rows (b, i) become output rows i*B + b
                step.acts = [torch.empty(n_dyn_acts, **f32) for _ in range(K)]
                z_all = torch.empty(K, N, R, **f32)
                l1_all, l2_all = torch.empty(K, N, Rs, **f32), torch.empty(K, N, Rs, **f32)
                for i in range(K):
                    check(lib.op3_dynamics_fwd(C.byref(d), C.byref(pss[i]), ptr(state.z), ptr(act), ptr(step.acts[i]),
                                               ptr(z_all[i]) if Rd else None, ptr(l1_all[i]), ptr(l2_all[i]), st), "dynamics_fwd")
                pick = lambda t, w: t.view(K, B, K, w).diagonal(dim1=0, dim2=2).permute(2, 0, 1).reshape(N, w)
                new.l1, new.l2 = pick(l1_all, Rs).contiguous(), pick(l2_all, Rs).contiguous()
                if Rd:
                    new.z[:, :Rd] = pick(z_all, R)[:, :Rd]
            check(lib.op3_sample_fwd(C.byref(d), N, ptr(new.l1), ptr(new.l2), ptr(new.eps), None, ptr(new.z), st), "sample")
            new.prior_kind = "alias"                                                  # op3_model.py:284-287
            step.dst = new
            return step

        steps, decodes = [], []
        cur, state = 0, s0
        cached = None
        if T1 and schedule[0] in (0, 2):
            if images is None:
                raise ValueError("refinement steps need images")
            cached = decode(s0, None, images[:, 0], 0.0, True, False)
            if need_grad:
                decodes.append(cached)
        for i, code in enumerate(schedule):
            eps_new = noise_fn(B, K, Rs, dev).reshape(N, Rs).contiguous()
            if code == 0:
                step = refine(state, cached, eps_new)
            elif code == 2:
                # next-step refinement (op3_model.py:368-376): refine on the current frame, then advance the frame index.
                # The action goes to the refinement net as add_fc_input, which the built 'size_dependent_conv'
                # architecture (added_fc_input_size=0) ignores (refinement_network.py:201-202).
                step = refine(state, cached, eps_new)
                cur += 1
            else:
                act = None
                if m.action_dim > 0:
                    if acts_t is None:
                        raise ValueError("this model has action_dim=%d: dynamics steps need actions" % m.action_dim)
                    act = acts_t[cur]
                step = dynamics(state, act, eps_new)
                cur += 1
            state = step.dst
            refine_after = i + 1 < T1 and schedule[i + 1] in (0, 2)
            img = None
            if lw[i] != 0.0 or refine_after:
                if images is None:
                    raise ValueError("loss / refinement at step %d needs images" % i)
                img = images[:, cur]            # IndexError past the last frame, like the reference (op3_model.py:383)
            elif i == T1 - 1 and images is not None:
                # only the final mse (against images[:, -1], op3_model.py:411) is needed: a rollout may run past the
                # last observed frame (mpc_stack.py:256-257: obs (B,1,...) with schedule [0,0,0,0,1]); the per-step
                # likelihood statistics of this decode carry zero weight, so any observed frame serves as their target
                img = images[:, min(cur, images.shape[1] - 1)]
            cached = decode(state, i, img, lw[i], refine_after, i == T1 - 1)
            if need_grad:
                steps.append(step)
                decodes.append(cached)
            else:                    # inference: nothing is kept for a backward pass
                step = None
                if not refine_after:
                    cached.acts = None

        sw = float(sum(lw))
        out = {"colors_list": colors_list.permute(1, 0, 2, 3, 4, 5), "masks_list": masks_list.permute(1, 0, 2, 3, 4, 5),
               "final_recon": final_recon}
        if sw == 0:
            out["total_loss"] = out["kle_loss"] = out["clog_prob"] = None
        else:
            w = self._loss_weights(lw, sw, dev)
            comb = (loss_rows[:T1, :3] * w[:, None]).sum(0)
            out["clog_prob"], out["kle_loss"], out["total_loss"] = comb[0], comb[1], comb[2]
        out["mse"] = loss_rows[T1 - 1, 3] if (T1 and images is not None) else torch.zeros((), **f32)
        out["state"] = state
        if tape is not None:
            tape.steps, tape.decodes, tape.s0, tape.from_params = steps, decodes, s0, from_params
            tape.B, tape.sw, tape.dims = B, sw, d
            tape.final_state = state
        return out, tape

    # ------------------------------------------------------------------ backward
    def backward(self, tape, g_total, g_kle, g_clog):
        """Accumulate d(g_total*total + g_kle*kle + g_clog*clog)/d params into a fresh flat gradient buffer and
        return it (same layout as flat_params())."""
        dev = self.flat_params().device
        # conv weight gradients on the library's weight-gradient lane (include/op3_b200.h, op3_set_wgrad_overlap): they fill
        # the SMs the small kernels of the dgrad chain leave idle.  Joined once per state below (d_dec and the scratch
        # buffers are reused from one state to the next) and before the gradients are finalised; the *_No_Sharing networks
        # reuse one scratch buffer for K calls per state and keep the serial order.
        overlap = bool(self.wgrad_overlap) and not (self.ns_dec or self.ns_ref)
        with torch.cuda.device(dev):
            if not overlap:
                return self._backward(tape, g_total, g_kle, g_clog, False)
            check(self.lib.op3_set_wgrad_overlap(1), "set_wgrad_overlap")
            try:
                return self._backward(tape, g_total, g_kle, g_clog, True)
            finally:
                self.lib.op3_wgrad_join(torch.cuda.current_stream(dev).cuda_stream)
                self.lib.op3_set_wgrad_overlap(0)

    wgrad_overlap = True

    def _backward(self, tape, g_total, g_kle, g_clog, overlap=False):
        m, lib = self.model, self.lib
        d, B = tape.dims, tape.B
        K, D, Rd, Rs = m.K, m.image_size, m.det_size, m.sto_size
        R, N, DD = Rd + Rs, B * K, D * D
        flat = self.flat_params()
        dev = flat.device
        st = torch.cuda.current_stream(dev).cuda_stream
        f32 = dict(device=dev, dtype=torch.float32)
        pss = self.params_structs()
        preps = self.prepared_all(st)
        ps, prep = pss[0], preps[0]
        gflat = torch.zeros_like(flat)
        gss = self._structs_for(gflat)
        gs = gss[0]
        pgs = [torch.zeros_like(p) for p in preps]
        pg = pgs[0]
        ns_dec, ns_ref, ns_dyn = self.ns_dec, self.ns_ref, self.ns_dyn
        d1 = Op3Dims(B=B, K=1, D=D, Rd=Rd, Rs=Rs, A=m.action_dim, beta=float(m.beta))
        dec_scratch = torch.empty(lib.op3_decoder_scratch_floats(C.byref(d), N), **f32)
        ref_scratch = None
        dyn_scratch = None
        d_dec = torch.empty(N, DD, 4, **f32)

        # gradient accumulators of every state of the schedule: ONE zero-filled slab (a fill kernel per tensor was 50 launches)
        per_state = N * (R + 2 * Rs + 2 * LSTM)
        n_states = len(tape.steps) + 1
        slab = torch.zeros(n_states, per_state, **f32)
        slab_next = [0]

        def grads_of(s):
            if s.dz is None:
                if slab_next[0] < n_states:
                    row = slab[slab_next[0]]
                    slab_next[0] += 1
                else:                                        # a prior owner outside the schedule's own states
                    row = torch.zeros(per_state, **f32)
                o = 0
                s.dz = row[o:o + N * R].view(N, R); o += N * R
                s.dl1 = row[o:o + N * Rs].view(N, Rs); o += N * Rs
                s.dl2 = row[o:o + N * Rs].view(N, Rs); o += N * Rs
                s.dh = row[o:o + N * LSTM].view(N, LSTM); o += N * LSTM
                s.dc = row[o:o + N * LSTM].view(N, LSTM)
            return s

        # decodes grouped by the state they decode
        dec_of = {id(dc.state): dc for dc in tape.decodes}
        d_in_for = {}   # id(decode) -> gradient of the refinement conv-input chunks
        states = [tape.s0] + [s.dst for s in tape.steps]
        for j in range(len(states) - 1, -1, -1):
            s = grads_of(states[j])
            dc = dec_of.get(id(s))
            w = dc.weight / tape.sw if (dc is not None and tape.sw) else 0.0
            w_clog = (g_total + g_clog) * w
            w_kl = (g_total * float(m.beta) + g_kle) * w / B
            d_in = d_in_for.pop(id(dc), None) if dc is not None else None
            if overlap:
                check(lib.op3_wgrad_join(st), "wgrad_join")
            if dc is not None and (w_clog != 0.0 or d_in is not None):
                check(lib.op3_mixture_bwd(C.byref(d), C.byref(dc.io), w_clog, ptr(d_in), ptr(d_dec), C.byref(gs), None, st),
                      "mixture_bwd")
                if not ns_dec:
                    check(lib.op3_decoder_bwd(C.byref(d), N, ptr(prep), ptr(s.z), ptr(dc.acts), ptr(d_dec), ptr(dec_scratch),
                                              ptr(s.dz), 1, ptr(pg), st), "decoder_bwd")
                else:
                    dd_v, dzp = d_dec.view(K, B, DD, 4), torch.empty(K, B, R, **f32)
                    for k in range(K):
                        check(lib.op3_decoder_bwd(C.byref(d1), B, ptr(preps[k]), ptr(dc.zp[k]), ptr(dc.acts[k]), ptr(dd_v[k]),
                                                  ptr(dec_scratch), ptr(dzp[k]), 0, ptr(pgs[k]), st), "decoder_bwd")
                    s.dz.view(B, K, R).add_(dzp.transpose(0, 1))
            # samples = eps*std(l2)+l1 and the KL term of this decode's loss
            owner = s.prior_owner() if w_kl != 0.0 else None
            if owner is not None:
                grads_of(owner)
            p1, p2 = s.prior_tensors()
            check(lib.op3_state_bwd(C.byref(d), N, ptr(s.l1), ptr(s.l2), ptr(s.eps), ptr(p1), ptr(p2), ptr(s.dz), w_kl,
                                    ptr(s.dl1), ptr(s.dl2), ptr(owner.dl1) if owner is not None else None,
                                    ptr(owner.dl2) if owner is not None else None, st), "state_bwd")
            if j == 0:
                break
            step = tape.steps[j - 1]
            prev = grads_of(step.src)
            if step.kind == "refine":
                if ref_scratch is None:
                    ref_scratch = torch.empty(lib.op3_refine_scratch_floats(C.byref(d)), **f32)
                if Rd:
                    prev.dz[:, :Rd] += s.dz[:, :Rd]          # deter_state is carried over (op3_model.py:257)
                if not ns_ref:
                    d_in_new = torch.empty(N, 5, DD, 4, **f32)
                    check(lib.op3_refine_bwd(C.byref(d), C.byref(ps), ptr(prep), C.byref(step.io), ptr(s.dl1), ptr(s.dl2),
                                             ptr(s.dh), ptr(s.dc), ptr(d_in_new), ptr(prev.dl1), ptr(prev.dl2), ptr(prev.dz),
                                             ptr(prev.dh), ptr(prev.dc), C.byref(gs), ptr(pg), ptr(ref_scratch), st), "refine_bwd")
                else:
                    d_in_p = torch.empty(K, B, 5, DD, 4, **f32)
                    tl1, tl2, tz = torch.zeros(K, B, Rs, **f32), torch.zeros(K, B, Rs, **f32), torch.zeros(K, B, R, **f32)
                    th, tc = torch.zeros(K, B, LSTM, **f32), torch.zeros(K, B, LSTM, **f32)
                    up = [t.view(K, B, -1) for t in (s.dl1, s.dl2, s.dh, s.dc)]          # output rows are slot-major
                    for k in range(K):
                        check(lib.op3_refine_bwd(C.byref(d1), C.byref(pss[k]), ptr(preps[k]), C.byref(step.io[k]), ptr(up[0][k]),
                                                 ptr(up[1][k]), ptr(up[2][k]), ptr(up[3][k]), ptr(d_in_p[k]), ptr(tl1[k]), ptr(tl2[k]),
                                                 ptr(tz[k]), ptr(th[k]), ptr(tc[k]), C.byref(gss[k]), ptr(pgs[k]), ptr(ref_scratch),
                                                 st), "refine_bwd")
                    d_in_new = d_in_p.transpose(0, 1).contiguous().view(N, 5, DD, 4)
                    for dst, t in ((prev.dl1, tl1), (prev.dl2, tl2), (prev.dz, tz), (prev.dh, th), (prev.dc, tc)):
                        dst.view(B, K, -1).add_(t.transpose(0, 1))
                d_in_for[id(step.dec)] = d_in_new
            else:
                if dyn_scratch is None:
                    dyn_scratch = torch.empty(lib.op3_dynamics_scratch_floats(C.byref(d)), **f32)
                dz_tmp = torch.empty(N, R, **f32)
                if not ns_dyn:
                    check(lib.op3_dynamics_bwd(C.byref(d), C.byref(ps), ptr(step.src.z), ptr(step.actions), ptr(step.acts),
                                               ptr(s.dz), ptr(s.dl1), ptr(s.dl2), ptr(dz_tmp), C.byref(gs), ptr(dyn_scratch), st),
                          "dynamics_bwd")
                    prev.dz += dz_tmp
                else:
                    # output row i*B + b came from row (b, i) of model i: that model sees this gradient there, zero elsewhere
                    def spread(t, w):
                        u = torch.zeros(K, B, K, w, **f32)
                        u.diagonal(dim1=0, dim2=2).copy_(t.view(K, B, w).permute(1, 2, 0))
                        return u.view(K, N, w)
                    uz, u1, u2 = spread(s.dz, R), spread(s.dl1, Rs), spread(s.dl2, Rs)
                    for i in range(K):
                        check(lib.op3_dynamics_bwd(C.byref(d), C.byref(pss[i]), ptr(step.src.z), ptr(step.actions), ptr(step.acts[i]),
                                                   ptr(uz[i]), ptr(u1[i]), ptr(u2[i]), ptr(dz_tmp), C.byref(gss[i]), ptr(dyn_scratch),
                                                   st), "dynamics_bwd")
                        prev.dz += dz_tmp
            s.dz = s.dl1 = s.dl2 = s.dh = s.dc = None
        # initial state parameters (op3_model.py:168-171)
        s0 = tape.s0
        if tape.from_params:
            named = dict(zip(self._names, self._offsets))
            def gview(name, n):
                o = named[name]
                return gflat[o:o + n]
            gview("initial_lambdas1", Rs).add_(s0.dl1.sum(0))
            gview("initial_lambdas2", Rs).add_(s0.dl2.sum(0))
            if Rd:
                gview("inital_deter_state", Rd).add_(s0.dz[:, :Rd].sum(0))
        if overlap:
            check(lib.op3_wgrad_join(st), "wgrad_join")
        for k in range(len(pgs)):        # per weight set; the shared networks' accumulators live in set 0 only
            check(lib.op3_finalize_grads(C.byref(d), ptr(pgs[k]), C.byref(gss[k]), st), "finalize_grads")
        s0.dz = s0.dl1 = s0.dl2 = s0.dh = s0.dc = None
        self.last_flat_grad = gflat
        return gflat

    def sub_images(self, colors, masks):
        """colors (B,K,3,D,D) * masks (B,K,1,D,D) (op3_model.py:612)."""
        colors, masks = colors.contiguous(), masks.contiguous()
        out = torch.empty_like(colors)
        n_slots = colors.shape[0] * colors.shape[1]
        st = torch.cuda.current_stream(colors.device).cuda_stream
        check(self.lib.op3_sub_images(n_slots, colors.shape[-1], ptr(colors), ptr(masks), ptr(out), st), "sub_images")
        return out

    def grad_views(self, gflat):
        return [gflat[o:o + p.numel()].view(p.shape) for (_, p), o in zip(self._named_params(), self._offsets)]
