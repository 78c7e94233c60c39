"""Mlp mirroring op3/torch/networks.py:18-74 (Linear+activation stack, `fcs` + `last_fc`).

The dynamics network runs all of its Mlps inside op3_dynamics_fwd/bwd (csrc/dynamics.cu); this class carries the
parameters (state_dict keys `fcs.N.weight/bias`, `last_fc.weight/bias`), the reference initialisation, and a standalone
inference-form `forward` (one op3_linear_fwd launch per layer, activation fused)."""
import torch
from torch import nn as nn
from torch.nn import functional as F

from op3_b200.torch import pytorch_util as ptu


def identity(x):
    return x


class Mlp(nn.Module):
    def __init__(self, hidden_sizes, output_size, input_size, init_w=3e-3, hidden_activation=F.relu,
                 output_activation=identity, hidden_init=ptu.fanin_init, b_init_value=0.1, layer_norm=False,
                 layer_norm_kwargs=None):
        super().__init__()
        if layer_norm:
            raise NotImplementedError("Mlp(layer_norm=True) is not used by OP3 and is not built")
        self.input_size, self.output_size = input_size, output_size
        self.hidden_activation, self.output_activation = hidden_activation, output_activation
        self.layer_norm = False
        self.fcs = nn.ModuleList()
        width = input_size
        for h in hidden_sizes:
            fc = nn.Linear(width, h)
            hidden_init(fc.weight)
            fc.bias.data.fill_(b_init_value)
            self.fcs.append(fc)
            width = h
        self.last_fc = nn.Linear(width, output_size)
        self.last_fc.weight.data.uniform_(-init_w, init_w)
        self.last_fc.bias.data.uniform_(-init_w, init_w)

    @staticmethod
    def _act_code(fn):
        name = getattr(fn, "__name__", type(fn).__name__).lower()
        codes = {"identity": 0, "elu": 1, "sigmoid": 2, "relu": 3, "tanh": 4}
        if name not in codes:
            raise NotImplementedError("Mlp activation %r has no fused kernel (identity/ELU/sigmoid/ReLU/tanh are built)" % name)
        return codes[name]

    def forward(self, input, return_preactivations=False):
        """networks.py:63-74 on CUDA tensors (detached, inference form; training goes through OP3Model.forward)."""
        from op3_b200 import _lib
        from op3_b200._lib import check, ptr
        if not input.is_cuda:
            raise RuntimeError("op3_b200: Mlp input is on %s; the OP3 kernels are CUDA-only (no CPU fallback)" % input.device)
        lib = _lib.load()
        lead = input.shape[:-1]
        h = input.detach().reshape(-1, input.shape[-1]).contiguous().float()
        if h.shape[1] != self.input_size:
            raise ValueError("Mlp expects %d input features, got %d" % (self.input_size, h.shape[1]))
        with torch.cuda.device(h.device):
            st = torch.cuda.current_stream(h.device).cuda_stream

            def linear(x, fc, act):
                w, b = fc.weight.detach().contiguous(), fc.bias.detach().contiguous()
                y = torch.empty(x.shape[0], fc.out_features, device=x.device)
                check(lib.op3_linear_fwd(x.shape[0], fc.in_features, fc.out_features, ptr(x), ptr(w), ptr(b), act, ptr(y), st),
                      "op3_linear_fwd")
                return y
            for fc in self.fcs:
                h = linear(h, fc, self._act_code(self.hidden_activation))
            code = self._act_code(self.output_activation)
            out = linear(h, self.last_fc, code)                 # output activation fused into the last layer
            if not return_preactivations:
                return out.reshape(*lead, -1)
            pre = out if code == 0 else linear(h, self.last_fc, 0)
        return out.reshape(*lead, -1), pre.reshape(*lead, -1)
