"""Mirror of op3/torch/op3_modules/op3_model.py: `create_model_v2` (:22-60) and `OP3Model` (:66-698) with the same
method names, argument meaning, return shapes, hidden-state dict schema and state_dict keys, executed by the
sm_100a kernels of libop3_b200.so through the schedule engine (op3_b200/engine.py).

Differences that are visible to a caller (all documented in INTEGRATION.md):
  * CUDA only -- CPU tensors / a CPU model raise RuntimeError (there is no fallback path);
  * autograd: `forward` / `run_schedule` return losses that back-propagate into the parameters (one fused backward
    through the engine); colors/masks/final_recon/mse/hidden-state outputs are returned detached;
  * `batch_internal_inference` does not chop the batch into chunks of 10 and does not decode a state twice.
"""
import numpy as np
import torch
from torch import nn
from torch.nn import Parameter

from op3_b200.engine import Engine, LSTM
from op3_b200.torch import pytorch_util as ptu
from op3_b200.torch.modules import LayerNorm2D, LayerNorm
from op3_b200.torch.op3_modules.decoder_network import DecoderNetwork, Decoder_Args, DecoderNetwork_v2_No_Sharing
from op3_b200.torch.op3_modules.refinement_network import (RefinementNetwork, Refinement_Args,
                                                            RefinementNetwork_v2_No_Sharing)
from op3_b200.torch.op3_modules.physics_network import PhysicsNetwork, Physics_Args, PhysicsNetwork_v2_No_Sharing


def create_model_v2(variant, det_size, sto_size, action_dim):
    """Same registry-driven factory as the reference (op3_model.py:22-60)."""
    K = variant["K"]
    kind, make = Refinement_Args[variant["refinement_model_type"]]
    if kind == "reg":
        refinement_net = RefinementNetwork(**make(sto_size))
    elif kind == "sequence_iodine":
        refinement_net = RefinementNetwork(**make(sto_size, action_dim))
    elif kind == "no_share":
        refinement_net = RefinementNetwork_v2_No_Sharing(**make(sto_size, K))
    else:
        raise ValueError("{}".format(kind))
    kind, make = Decoder_Args[variant["decoder_model_type"]]
    if kind == "reg":
        decoder_net = DecoderNetwork(**make(det_size + sto_size))
    elif kind == "reg_no_share":
        decoder_net = DecoderNetwork_v2_No_Sharing(**make(det_size + sto_size, K))
    else:
        raise ValueError("{}".format(kind))
    kind, make = Physics_Args[variant["dynamics_model_type"]]
    if kind == "reg":
        dynamics_net = PhysicsNetwork(**make(det_size, sto_size, action_dim))
    elif kind == "reg_no_share":
        dynamics_net = PhysicsNetwork_v2_No_Sharing(**make(det_size, sto_size, action_dim, K))
    else:
        raise ValueError("{}".format(variant["dynamics_model_type"][0]))
    model = OP3Model(refinement_net, dynamics_net, decoder_net, det_size, sto_size, **variant["extra_args"])
    model.set_k(K)
    return model


class _ScheduleFn(torch.autograd.Function):
    """Autograd hand-over: one node for the whole unrolled schedule; backward = Engine.backward."""

    @staticmethod
    def forward(ctx, model, images, actions, init_state, schedule, loss_schedule, *params):
        out, tape = model.engine.forward(images, actions, init_state, schedule, loss_schedule, model._noise, True)
        ctx.model, ctx.tape = model, tape
        B = tape.B
        state = model._state_to_tensors(out["state"], B)
        nondiff = [out["colors_list"], out["masks_list"], out["final_recon"], out["mse"]] + [t for t in state if t is not None]
        ctx.mark_non_differentiable(*nondiff)
        return (out["total_loss"], out["kle_loss"], out["clog_prob"], out["colors_list"], out["masks_list"],
                out["final_recon"], out["mse"]) + tuple(state)

    @staticmethod
    def backward(ctx, g_total, g_kle, g_clog, *unused):
        f = lambda g: 0.0 if g is None else float(g)
        eng = ctx.model.engine
        gflat = eng.backward(ctx.tape, f(g_total), f(g_kle), f(g_clog))
        return (None,) * 6 + tuple(eng.grad_views(gflat))


class OP3Model(nn.Module):
    def __init__(self, refine_net, dynamics_net, decode_net, det_size, sto_size, beta=1, deterministic_sampling=False):
        super().__init__()
        self.refinement_net, self.dynamics_net, self.decode_net = refine_net, dynamics_net, decode_net
        self.K = None
        self.det_size, self.sto_size = det_size, sto_size
        self.full_rep_size = det_size + sto_size
        self.deterministic_sampling = deterministic_sampling
        self.sigma = 0.1
        self.set_beta(beta)
        self.layer_norms_2d = nn.ModuleList([LayerNorm2D(c) for c in [1, 1, 1, 3]])
        self.layer_norms_1d = nn.ModuleList([LayerNorm(sto_size, center=True, scale=True) for _ in range(2)])
        self.eval_mode = False
        if det_size != 0:
            self.inital_deter_state = Parameter(ptu.randn((det_size)) / np.sqrt(det_size))   # (sic) key name kept
        self.initial_lambdas1 = Parameter(ptu.randn((sto_size)) / np.sqrt(sto_size))
        self.initial_lambdas2 = Parameter(ptu.randn((sto_size)) / np.sqrt(sto_size))
        self.debug = {}
        # properties of the sub-networks the kernels need
        self.action_dim = dynamics_net.action_size
        self.image_size = decode_net.input_width
        if refine_net.input_width != self.image_size:
            raise ValueError("decoder and refinement networks disagree on the image size")
        self.inference_chunk = 512          # samples per engine call in batch_internal_inference
        self._engine = None

    # ------------------------------------------------------------------ plumbing
    @property
    def engine(self):
        if self._engine is None:
            object.__setattr__(self, "_engine", Engine(self))
        return self._engine

    def __getstate__(self):
        # the engine holds a ctypes handle and device buffers keyed to THIS module: never pickled / deep-copied
        state = self.__dict__.copy()
        state["_engine"] = None
        return state

    def _replicate_for_data_parallel(self):
        # nn.DataParallel (train_op3.py:99-100 with variant['dataparallel']) shallow-copies __dict__, so every replica
        # would share one Engine bound to device 0's parameters.  Multi-GPU here is one process per GPU (torchrun).
        if torch.cuda.device_count() > 1:
            raise RuntimeError("op3_b200: nn.DataParallel over several GPUs is not supported; launch one process per GPU "
                               "with torchrun (OP3Trainer all-reduces the gradients) and set variant['dataparallel']=False")
        return super()._replicate_for_data_parallel()

    def _noise(self, B, K, Rs, dev):
        if self.deterministic_sampling:          # samples = means (op3_model.py:148-149)
            return torch.zeros(B, K, Rs, device=dev)
        if torch.cuda.is_current_stream_capturing():
            # inside a CUDA-graph capture (OP3Trainer.use_cuda_graph) the draw must be a device-side philox kernel:
            # every replay then gets fresh noise from the registered generator state
            return torch.randn(B, K, Rs, device=dev)
        # drawn on the inputs' device even when ptu.set_gpu_mode was never called: a CPU draw would cost a pageable
        # host->device copy (= a host sync) per schedule step
        return ptu.randn(B, K, Rs, torch_device=dev).to(dev)

    def _state_to_tensors(self, s, B):
        K, Rd, Rs = self.K, self.det_size, self.sto_size
        p1, p2 = s.prior_tensors()
        deter = s.z[:, :Rd].reshape(B, K, Rd) if Rd else None
        return [p1.view(B, K, Rs), p2.view(B, K, Rs), deter, s.l1.view(B, K, Rs), s.l2.view(B, K, Rs),
                s.h.view(B, K, LSTM), s.c.view(B, K, LSTM), s.z[:, Rd:].reshape(B, K, Rs)]

    @staticmethod
    def _tensors_to_state(t):
        return {"prior": {"lambdas1": t[0], "lambdas2": t[1]},
                "post": {"deter_state": t[2], "lambdas1": t[3], "lambdas2": t[4], "extra_info": [t[5], t[6]],
                         "samples": t[7]}}

    def set_k(self, k):
        self.K = k
        self.dynamics_net.set_k(k)

    def set_beta(self, beta):
        if self.deterministic_sampling and beta != 0:
            print("DANGER ALERT! Having beta set to {} is mathematically incorrect when sampling deterministically".format(beta))
        self.beta = beta

    def _flatten_first_two(self, x):
        return x if x is None else x.reshape([x.shape[0] * x.shape[1]] + list(x.shape[2:]))

    def _unflatten_first(self, x, k):
        return x if x is None else x.reshape([-1, k] + list(x.shape[1:]))

    def get_full_tensor_state(self, hidden_state):
        if self.det_size == 0:
            return hidden_state["post"]["samples"]
        return torch.cat([hidden_state["post"]["deter_state"], hidden_state["post"]["samples"]], dim=2)

    # ------------------------------------------------------------------ core
    def run_schedule(self, images, actions, initial_hidden_state, schedule, loss_schedule, should_detach=False):
        """op3_model.py:333-434.  Returns the same 8-tuple: colors_list (B,T1,K,3,D,D), masks_list (B,T1,K,1,D,D),
        final_recon (B,3,D,D), total_loss, total_kle_loss, total_clog_prob, mse, hidden_state."""
        self.debug["schedule"] = schedule
        if images is None and actions is None:
            raise ValueError("Need either images or actions")
        schedule = np.asarray(schedule)
        loss_schedule = np.asarray(loss_schedule)
        B = images.shape[0] if images is not None else actions.shape[0]
        params = list(self.parameters())
        need_grad = (torch.is_grad_enabled() and not should_detach and float(np.sum(loss_schedule)) != 0.0
                     and any(p.requires_grad for p in params))
        if need_grad:
            res = _ScheduleFn.apply(self, images, actions, initial_hidden_state, schedule, loss_schedule, *params)
            total, kle, clog, colors, masks, recon, mse = res[:7]
            state_t = list(res[7:])
        else:
            out, _ = self.engine.forward(images, actions, initial_hidden_state, schedule, loss_schedule, self._noise, False)
            total, kle, clog = out["total_loss"], out["kle_loss"], out["clog_prob"]
            colors, masks, recon, mse = out["colors_list"], out["masks_list"], out["final_recon"], out["mse"]
            state_t = self._state_to_tensors(out["state"], B)
        return colors, masks, recon, total, kle, clog, mse, self._tensors_to_state(state_t)

    def forward(self, images, actions, initial_hidden_state, schedule, loss_schedule):
        return self.run_schedule(images, actions, initial_hidden_state, schedule, loss_schedule)

    # ------------------------------------------------------------------ single steps (inference helpers)
    def _step_from(self, hidden_states, images, actions, code):
        imgs = None if images is None else images.unsqueeze(1)
        acts = None if actions is None else actions.unsqueeze(1)
        out, _ = self.engine.forward(imgs, acts, hidden_states, [code], [0], self._noise, False)
        B = hidden_states["post"]["samples"].shape[0]
        return self._tensors_to_state(self._state_to_tensors(out["state"], B))

    def refine(self, hidden_states, images, action=None, previous_decode_loss_info=None):
        """op3_model.py:195-264 (inference form): one refinement step on `images` (B,3,D,D).  The gradient inputs of
        the refinement network are closed forms computed by the kernels, so -- unlike the reference, whose
        autograd.grad needs lambdas that require grad -- any state (detached or not) can be refined."""
        # `action` is only forwarded to the refinement net as add_fc_input (op3_model.py:233-244); the built
        # 'size_dependent_conv' architecture has added_fc_input_size=0 and ignores it (refinement_network.py:201-202)
        return self._step_from(hidden_states, images, None, 0)

    def dynamics(self, hidden_states, actions):
        """op3_model.py:269-296 (inference form): one dynamics step with `actions` (B,A) or None."""
        return self._step_from(hidden_states, None, actions, 1)

    def decode(self, hidden_states):
        """op3_model.py:300-308: colors (B,K,3,D,D), mask (B,K,1,D,D), mask_logits (B,K,1,D,D)."""
        from op3_b200.torch.op3_modules._standalone import decoder_forward
        B, K = hidden_states["post"]["samples"].shape[:2]
        z = self._flatten_first_two(self.get_full_tensor_state(hidden_states))
        if getattr(self.decode_net, "no_sharing", False):
            logits, colors = self.decode_net(z.contiguous())             # one decoder per slot, rows slot-major
        else:
            logits, colors = decoder_forward(self.decode_net, z, engine=self.engine)
        D = self.image_size
        logits = logits.view(B, K, 1, D, D)
        return colors.view(B, K, 3, D, D), torch.softmax(logits, dim=1), logits

    def get_loss(self, hidden_states, colors, mask, target_imgs):
        """op3_model.py:313-327 as a standalone (inference-form, detached) call: colors (B,K,3,D,D), mask (B,K,1,D,D),
        target_imgs (B,3,D,D) -> (color_probs (B,K,D,D), pixel_complete_log_likelihood (B,D,D), kle_loss,
        complete_log_likelihood, total_loss).  Training uses the fused form inside run_schedule (op3_mixture_fwd)."""
        import ctypes as C
        from op3_b200 import _lib
        from op3_b200._lib import check, ptr
        lib = _lib.load()
        for name, t in [("colors", colors), ("mask", mask), ("target_imgs", target_imgs)]:
            if not t.is_cuda:
                raise RuntimeError("op3_b200: %s is on %s; the OP3 kernels are CUDA-only (no CPU fallback)" % (name, t.device))
        B, K, _, D, _ = colors.shape
        dev = colors.device
        d = self.engine.dims(B)
        if K != d.K or D != d.D:
            raise ValueError("colors must be (B,%d,3,%d,%d), got %s" % (d.K, d.D, d.D, tuple(colors.shape)))
        f = lambda t, shape: t.detach().reshape(shape).contiguous().float()
        post, prior = hidden_states["post"], hidden_states["prior"]
        Rs = self.sto_size
        tens = [f(colors, (B, K, 3, D, D)), f(mask, (B, K, 1, D, D)), f(target_imgs, (B, 3, D, D)),
                f(post["lambdas1"], (B * K, Rs)), f(post["lambdas2"], (B * K, Rs)),
                f(prior["lambdas1"].to(dev).expand(B, K, Rs), (B * K, Rs)), f(prior["lambdas2"].to(dev).expand(B, K, Rs), (B * K, Rs))]
        color_probs = torch.empty(B, K, D, D, device=dev)
        pixel_ll = torch.empty(B, D, D, device=dev)
        losses = torch.empty(3, device=dev)
        scratch = torch.empty(lib.op3_get_loss_scratch_floats(C.byref(d)) // 2 + 1, device=dev, dtype=torch.float64)
        st = torch.cuda.current_stream(dev).cuda_stream
        check(lib.op3_get_loss(C.byref(d), *[ptr(t) for t in tens], ptr(color_probs), ptr(pixel_ll), ptr(losses),
                               scratch.data_ptr(), st), "op3_get_loss")
        return color_probs, pixel_ll, losses[0], losses[1], losses[2]

    # ------------------------------------------------------------------ state utilities (op3_model.py:446-558)
    def get_rprp_schedule(self, seed_steps, num_images, num_refine_per_physics):
        schedule = np.zeros(seed_steps + (num_images - 1) * (num_refine_per_physics + 1))
        schedule[seed_steps::(num_refine_per_physics + 1)] = 1
        return schedule

    @staticmethod
    def _map_state(state, fn, prior2_from_prior1=False):
        post, prior = state["post"], state["prior"]
        return {"prior": {"lambdas1": fn(prior["lambdas1"]),
                          "lambdas2": fn(prior["lambdas1"] if prior2_from_prior1 else prior["lambdas2"])},
                "post": {"deter_state": None if post["deter_state"] is None else fn(post["deter_state"]),
                         "lambdas1": fn(post["lambdas1"]), "lambdas2": fn(post["lambdas2"]),
                         "extra_info": [fn(post["extra_info"][0]), fn(post["extra_info"][1])],
                         "samples": fn(post["samples"])}}

    def detach_state(self, cur_hidden_state):
        return self._map_state(cur_hidden_state, lambda t: t.detach())

    def replicate_state(self, cur_hidden_state, n):
        """B=1 -> B=n.  Keeps the reference quirk prior.lambdas2 := prior.lambdas1 (op3_model.py:483-484, trap T7)."""
        if cur_hidden_state is None:
            return None
        return self._map_state(cur_hidden_state, lambda t: t.repeat(n, 1, 1), prior2_from_prior1=True)

    def select_specific_state(self, cur_hidden_state, indxs):
        """Same quirk as replicate_state (op3_model.py:502-503)."""
        return self._map_state(cur_hidden_state, lambda t: t[indxs], prior2_from_prior1=True)

    def _stack_state(self, array_of_states):
        cat = lambda xs: None if xs[0] is None else torch.cat(xs)
        g = lambda f: [f(s) for s in array_of_states]
        return {"prior": {"lambdas1": cat(g(lambda s: s["prior"]["lambdas1"])), "lambdas2": cat(g(lambda s: s["prior"]["lambdas2"]))},
                "post": {"deter_state": cat(g(lambda s: s["post"]["deter_state"])),
                         "lambdas1": cat(g(lambda s: s["post"]["lambdas1"])), "lambdas2": cat(g(lambda s: s["post"]["lambdas2"])),
                         "extra_info": [cat(g(lambda s: s["post"]["extra_info"][0])), cat(g(lambda s: s["post"]["extra_info"][1]))],
                         "samples": cat(g(lambda s: s["post"]["samples"]))}}

    # ------------------------------------------------------------------ batched inference / MPC (op3_model.py:569-698)
    def batch_internal_inference(self, obs, actions, initial_hidden_state, schedule, select_best=False, figure_path=None,
                                 batch_size=10):
        """Same inputs/outputs as the reference.  `batch_size` (a memory work-around there: chunks of 10) only acts as
        a lower bound here; the engine processes up to `self.inference_chunk` samples per call."""
        schedule = np.asarray(schedule)
        loss_schedule = np.zeros_like(schedule)
        if obs is not None:
            B = obs.shape[0]
        elif actions is not None:
            B = actions.shape[0]
        else:
            raise ValueError("Unknown size of inputs!")
        chunk = max(int(batch_size), int(self.inference_chunk))
        info = {"colors": [], "masks": [], "sub_images": [], "final_recon": [], "state": []}
        for start in range(0, B, chunk):
            end = min(start + chunk, B)
            b_obs = None if obs is None else obs[start:end]
            b_act = None if actions is None else actions[start:end]
            if initial_hidden_state is None or initial_hidden_state["post"]["samples"].shape[0] == 1:
                b_state = self.replicate_state(initial_hidden_state, end - start)
            else:
                b_state = self.select_specific_state(initial_hidden_state, slice(start, end))
            colors, masks, final_recon, _, _, _, _, state = self.run_schedule(b_obs, b_act, b_state, schedule, loss_schedule,
                                                                              should_detach=True)
            c_last, m_last = colors[:, -1], masks[:, -1]
            info["colors"].append(c_last)
            info["masks"].append(m_last)
            info["sub_images"].append(self.engine.sub_images(c_last, m_last))
            info["final_recon"].append(final_recon)
            info["state"].append(state)
        for k in ["colors", "masks", "sub_images", "final_recon"]:
            info[k] = info[k][0] if len(info[k]) == 1 else torch.cat(info[k])
        info["state"] = info["state"][0] if len(info["state"]) == 1 else self._stack_state(info["state"])
        if select_best:
            mses = torch.pow(info["final_recon"] - obs[:, -1], 2).mean((-1, -2, -3))
            best = int(torch.argmin(mses))
            for k in info:
                if k == "state":
                    info[k] = self.select_specific_state(info[k], [best])
                else:
                    info[k] = info[k][best:best + 1]
        return info

    def get_activation_values(self, initial_hidden_state, actions, batch_size=10):
        """op3_model.py:645-664: state-action attention (B,K) of B probe actions on a B=1 state.  Reference quirk kept:
        within each chunk of `batch_size` actions the rows are paired with `actions[chunk].repeat(K, 1)` (TILED, not
        interleaved per slot), so state row (i, k) meets action (i*K + k) mod b of its chunk (:661) -- the attention is
        a per-row quantity, so it is evaluated here with one action per row (K=1 rows of the dynamics kernels)."""
        from op3_b200.torch.op3_modules._standalone import physics_forward
        B, K = actions.shape[0], self.K
        state = self.replicate_state(initial_hidden_state, B)
        z = self._flatten_first_two(self.get_full_tensor_state(state)).contiguous()          # (B*K, R)
        rows = []
        for start in range(0, B, batch_size):
            b = min(batch_size, B - start)
            rows.append(actions[start:start + b].repeat(K, 1))                                 # (b*K, A), tiled
        acts = torch.cat(rows)
        act_att = physics_forward(self.dynamics_net, z, acts, K=1, engine=self.engine)[3]    # (B*K, 1)
        return act_att.view(B, K).detach()

    def get_all_activation_values(self, initial_hidden_state, actions, batch_size=10):
        """op3_model.py:667-698: (state_action_attention (B,K), interaction_attention (B,K,K-1), |d enc| (B,K),
        |d lambda| (B,K)) for B probe actions applied to a B=1 state."""
        from op3_b200.torch.op3_modules._standalone import physics_forward
        B, K = actions.shape[0], self.K
        state = self.replicate_state(initial_hidden_state, B)
        z = self._flatten_first_two(self.get_full_tensor_state(state)).contiguous()
        acts = self._flatten_first_two(actions.unsqueeze(1).repeat(1, K, 1))
        deter, l1, l2, act_att, pair_att, deltas = physics_forward(self.dynamics_net, z, acts, K=K, engine=self.engine)
        lam_d = torch.abs(self._flatten_first_two(state["post"]["lambdas1"]) - l1).sum(1)
        if deter is not None:
            lam_d = lam_d + torch.abs(z[:, :self.det_size] - deter).sum(1)
        return (act_att.view(B, K).detach(), pair_att.view(B, K, K - 1).detach(), deltas.view(B, K).detach(),
                lam_d.view(B, K).detach())
