"""ctypes binding of libop3_b200.so (the C ABI declared in include/op3_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OP3_B200_LIB") or os.path.join(_HERE, "libop3_b200.so")   # env override: experiment builds (tools/)

c_fp = C.c_void_p  # device pointers are passed as integers


class Op3Dims(C.Structure):
    _fields_ = [("B", C.c_int), ("K", C.c_int), ("D", C.c_int), ("Rd", C.c_int), ("Rs", C.c_int), ("A", C.c_int),
                ("beta", C.c_float)]


class Op3Linear(C.Structure):
    _fields_ = [("w", c_fp), ("b", c_fp)]


class Op3Mlp(C.Structure):
    _fields_ = [("fc0", Op3Linear), ("last", Op3Linear)]


class Op3Params(C.Structure):
    _fields_ = [
        ("inital_deter_state", c_fp), ("initial_lambdas1", c_fp), ("initial_lambdas2", c_fp),
        ("ref_conv_w", c_fp * 3), ("ref_conv_b", c_fp * 3), ("ref_fc", Op3Linear),
        ("lstm_w_ih", c_fp), ("lstm_w_hh", c_fp), ("lstm_b_ih", c_fp), ("lstm_b_hh", c_fp),
        ("ref_last_fc", Op3Linear), ("ref_last_fc2", Op3Linear),
        ("dyn_inertia", Op3Mlp), ("dyn_action_enc", Op3Mlp), ("dyn_action_effect", Op3Mlp), ("dyn_action_att", Op3Mlp),
        ("dyn_pairwise", Op3Mlp), ("dyn_inter_effect", Op3Mlp), ("dyn_inter_att", Op3Mlp), ("dyn_merge", Op3Mlp),
        ("dyn_det_out", Op3Mlp), ("dyn_lam1_out", Op3Mlp), ("dyn_lam2_out", Op3Mlp),
        ("dec_conv_w", c_fp * 5), ("dec_conv_b", c_fp * 5),
        ("ln2d_gamma", c_fp * 4), ("ln2d_beta", c_fp * 4), ("ln1d_scale", c_fp * 2), ("ln1d_center", c_fp * 2),
    ]


class Op3MixtureIO(C.Structure):
    _fields_ = [
        ("dec_out", c_fp), ("image", c_fp), ("image_bstride", C.c_longlong), ("mse_image", c_fp),
        ("mse_image_bstride", C.c_longlong), ("post_l1", c_fp), ("post_l2", c_fp), ("prior_l1", c_fp),
        ("prior_l2", c_fp), ("mask", c_fp), ("colors", c_fp), ("final_recon", c_fp), ("stats", c_fp), ("losses", c_fp),
        ("mix1", c_fp), ("mix2", c_fp), ("smp", c_fp), ("seed", c_fp), ("sync", c_fp),
    ]


class Op3RefineIO(C.Structure):
    _fields_ = [
        ("dec_out", c_fp), ("mix1", c_fp), ("mix2", c_fp), ("smp", c_fp), ("dz", c_fp), ("z", c_fp), ("l1", c_fp),
        ("l2", c_fp), ("eps", c_fp), ("prior1", c_fp), ("prior2", c_fp), ("prior_is_post", C.c_int), ("h", c_fp),
        ("c", c_fp), ("acts", c_fp), ("l1_out", c_fp), ("l2_out", c_fp), ("h_out", c_fp), ("c_out", c_fp),
    ]


# name -> (restype, argtypes); used both to bind and by tests to check that every symbol is exported
_P = C.POINTER
SIGNATURES = {
    "op3_version": (C.c_char_p, []),
    "op3_last_error": (C.c_char_p, []),
    "op3_launch_count": (C.c_ulonglong, []),
    "op3_set_precision": (C.c_int, [C.c_int]),
    "op3_get_precision": (C.c_int, []),
    "op3_set_fused_mlp": (C.c_int, [C.c_int]),
    "op3_set_tcq_pair": (C.c_int, [C.c_int]),
    "op3_set_wgrad_overlap": (C.c_int, [C.c_int]),
    "op3_wgrad_join": (C.c_int, [C.c_void_p]),
    "op3_profile_enable": (C.c_int, [C.c_int]),
    "op3_profile_collect": (C.c_int, [_P(C.c_double), _P(C.c_double), _P(C.c_longlong)]),
    "op3_prepared_floats": (C.c_size_t, [_P(Op3Dims)]),
    "op3_prepare": (C.c_int, [_P(Op3Dims), _P(Op3Params), c_fp, c_fp]),
    "op3_finalize_grads": (C.c_int, [_P(Op3Dims), c_fp, _P(Op3Params), c_fp]),
    "op3_decoder_acts_floats": (C.c_size_t, [_P(Op3Dims), C.c_int]),
    "op3_decoder_scratch_floats": (C.c_size_t, [_P(Op3Dims), C.c_int]),
    "op3_decoder_fwd": (C.c_int, [_P(Op3Dims), C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "op3_decoder_bwd": (C.c_int, [_P(Op3Dims), C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, C.c_int, c_fp, c_fp]),
    "op3_mixture_stats_floats": (C.c_size_t, [_P(Op3Dims)]),
    "op3_mixture_bwd_scratch_floats": (C.c_size_t, [_P(Op3Dims)]),
    "op3_mixture_fwd": (C.c_int, [_P(Op3Dims), _P(Op3Params), _P(Op3MixtureIO), C.c_int, c_fp]),
    "op3_mixture_bwd": (C.c_int, [_P(Op3Dims), _P(Op3MixtureIO), C.c_float, c_fp, c_fp, _P(Op3Params), c_fp, c_fp]),
    "op3_sample_fwd": (C.c_int, [_P(Op3Dims), C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "op3_state_bwd": (C.c_int, [_P(Op3Dims), C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, C.c_float, c_fp, c_fp, c_fp,
                                c_fp, c_fp]),
    "op3_refine_acts_floats": (C.c_size_t, [_P(Op3Dims)]),
    "op3_refine_scratch_floats": (C.c_size_t, [_P(Op3Dims)]),
    "op3_refine_fwd": (C.c_int, [_P(Op3Dims), _P(Op3Params), c_fp, _P(Op3RefineIO), c_fp]),
    "op3_refine_net_fwd": (C.c_int, [_P(Op3Dims), _P(Op3Params), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                     c_fp]),
    "op3_refine_bwd": (C.c_int, [_P(Op3Dims), _P(Op3Params), c_fp, _P(Op3RefineIO), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                 c_fp, c_fp, c_fp, c_fp, _P(Op3Params), c_fp, c_fp, c_fp]),
    "op3_dynamics_acts_floats": (C.c_size_t, [_P(Op3Dims)]),
    "op3_dynamics_scratch_floats": (C.c_size_t, [_P(Op3Dims)]),
    "op3_dynamics_fwd": (C.c_int, [_P(Op3Dims), _P(Op3Params), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "op3_dynamics_probes": (C.c_int, [_P(Op3Dims), c_fp, c_fp, c_fp, c_fp, c_fp]),
    "op3_dynamics_bwd": (C.c_int, [_P(Op3Dims), _P(Op3Params), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, _P(Op3Params),
                                   c_fp, c_fp]),
    "op3_adam_clip": (C.c_int, [C.c_longlong, c_fp, c_fp, c_fp, c_fp, C.c_int, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_float, c_fp, c_fp]),
    "op3_mpc_cost": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "op3_mpc_cost_ex": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "op3_get_loss_scratch_floats": (C.c_size_t, [_P(Op3Dims)]),
    "op3_get_loss": (C.c_int, [_P(Op3Dims), c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "op3_sub_images": (C.c_int, [C.c_longlong, C.c_int, c_fp, c_fp, c_fp, c_fp]),
    "op3_prepare_frames": (C.c_int, [C.c_longlong, C.c_int, C.c_int, c_fp, c_fp, c_fp]),
    "op3_linear_fwd": (C.c_int, [C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, C.c_int, c_fp, c_fp]),
    "op3_cem_refit": (C.c_int, [C.c_int, C.c_int, C.c_int, c_fp, c_fp, C.c_float, c_fp, c_fp, c_fp]),
    "op3_cem_resample": (C.c_int, [C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
}

_lib = None


def load():
    """Load (once) and return the bound library.  Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "op3_b200: %s is missing. Build it with `python -m op3_b200._build` (nvcc, sm_100a). "
            "There is no CPU / PyTorch fallback for the OP3 hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().op3_last_error().decode("utf-8", "replace")
        raise RuntimeError("op3_b200 %s failed (code %d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL).  CUDA fp32 contiguous tensors only."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("op3_b200: expected a CUDA tensor, got %s (there is no CPU path)" % t.device)
    return t.data_ptr()
