"""Mlp parameter container mirroring op3/torch/networks.py:18-74 (Linear+activation stack, `fcs` + `last_fc`).

The dynamics network runs all of its Mlps inside op3_dynamics_fwd/bwd (csrc/dynamics.cu); this class carries the
parameters (state_dict keys `fcs.N.weight/bias`, `last_fc.weight/bias`) and the reference initialisation."""
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

    def forward(self, input, return_preactivations=False):
        raise NotImplementedError("Mlp is executed inside the fused dynamics kernels (PhysicsNetwork.forward); "
                                  "a standalone Mlp.forward is not part of the OP3 path")
