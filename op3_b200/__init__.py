"""op3_b200: B200-native (sm_100a) kernels behind the OP3 iterative amortized-inference step.

Package layout
  csrc/                    hand-written CUDA kernels + the C ABI (include/op3_b200.h)
  _lib.py / _build.py      ctypes binding and in-tree nvcc build of libop3_b200.so
  engine.py                schedule executor (forward tape + manual backward) over the C ABI
  torch/                   mirror of the reference's op3/torch module surface (same names, kwargs, state_dict keys)
  parallel.py              one-process-per-GPU data parallelism (NCCL gradient allreduce, MPC candidate sharding)
"""
__version__ = "0.1.0"

_ALIASES = {
    "op3.torch.pytorch_util": "op3_b200.torch.pytorch_util",
    "op3.torch.modules": "op3_b200.torch.modules",
    "op3.torch.networks": "op3_b200.torch.networks",
    "op3.torch.conv_networks": "op3_b200.torch.conv_networks",
    "op3.torch.op3_modules.op3_model": "op3_b200.torch.op3_modules.op3_model",
    "op3.torch.op3_modules.op3_trainer": "op3_b200.torch.op3_modules.op3_trainer",
    "op3.torch.op3_modules.decoder_network": "op3_b200.torch.op3_modules.decoder_network",
    "op3.torch.op3_modules.refinement_network": "op3_b200.torch.op3_modules.refinement_network",
    "op3.torch.op3_modules.physics_network": "op3_b200.torch.op3_modules.physics_network",
    "op3.torch.data_management.dataset": "op3_b200.torch.data_management.dataset",
}


PRECISIONS = {"fp32": 0, "tf32x3": 1, "tf32": 2, "f16x3": 3}


def set_precision(name):
    """Arithmetic of the 5x5 convolutions: 'fp32' (CUDA cores, exact), 'tf32x3' (tcgen05, fp32-grade operand
    splitting), 'tf32' (tcgen05, 1xTF32 speed mode) or 'f16x3' (tcgen05, fp16 hi/lo split with per-image operand
    scaling for forward/dgrad: fp32-grade like tf32x3 at half the operand bytes; weight gradients stay 3xTF32).  Process-wide; derived weights are rebuilt on the next call."""
    from . import _lib
    _lib.check(_lib.load().op3_set_precision(PRECISIONS[name]), "op3_set_precision")


def get_precision():
    from . import _lib
    mode = _lib.load().op3_get_precision()
    return [k for k, v in PRECISIONS.items() if v == mode][0]


def install():
    """Make `import op3.torch.op3_modules.op3_model` (etc.) resolve to this package so that the reference's
    exps/train_op3/train_op3.py and MPC scripts run unchanged on top of the sm_100a kernels.  Call before the
    reference scripts import their modules."""
    import importlib
    import sys
    import types
    for ref_name, ours in _ALIASES.items():
        mod = importlib.import_module(ours)
        sys.modules[ref_name] = mod
        # `import op3.torch.op3_modules.op3_model as m` walks attributes from the top-level package, so the alias must
        # also hang off its parent packages (the reference's own, when its checkout is importable; bare namespace
        # stand-ins otherwise)
        parent_name, leaf = ref_name.rsplit(".", 1)
        parts = parent_name.split(".")
        for i in range(1, len(parts) + 1):
            name = ".".join(parts[:i])
            if name not in sys.modules:
                try:
                    importlib.import_module(name)
                except ImportError:
                    pkg = types.ModuleType(name)
                    pkg.__path__ = []
                    sys.modules[name] = pkg
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], sys.modules[name])
        setattr(sys.modules[parent_name], leaf, mod)
