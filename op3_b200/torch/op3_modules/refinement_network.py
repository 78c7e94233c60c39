"""Mirror of op3/torch/op3_modules/refinement_network.py: registry `Refinement_Args` (:14-80) and RefinementNetwork
(:83-231).  The conv stack, pooling, FC, LSTM cell and heads run in op3_refine_fwd/bwd (csrc/refine.cu)."""
import numpy as np
import torch
from torch import nn

from op3_b200.torch.networks import identity
from op3_b200.torch import pytorch_util as ptu
from op3_b200.torch.pytorch_util import from_numpy


def _size_dependent(repsize, **extra):
    return dict(input_width=64, input_height=64, input_channels=17, paddings=[2, 2, 2], kernel_sizes=[5, 5, 5],
                n_channels=[32, 32, repsize], strides=[1, 1, 1], hidden_sizes=[128], output_size=repsize, lstm_size=128,
                lstm_input_size=repsize * 5 + 128, added_fc_input_size=0, hidden_activation=nn.ELU(), **extra)


Refinement_Args = dict(
    large_reg=("reg", lambda repsize: dict(_size_dependent(repsize), paddings=[0, 0, 0], n_channels=[32, 32, 32])),
    size_dependent_conv=("reg", lambda repsize: _size_dependent(repsize)),
    large_sequence=("sequence_iodine", lambda repsize, action_size: dict(
        input_width=64, input_height=64, input_channels=17, paddings=[0] * 4, kernel_sizes=[5] * 4,
        n_channels=[64] * 4, strides=[2] * 4, hidden_sizes=[128, 128], output_size=repsize, lstm_size=256,
        lstm_input_size=repsize * 5 + 128, added_fc_input_size=action_size, hidden_activation=nn.ELU())),
    size_dependent_conv_no_share=("no_share", lambda repsize, k: _size_dependent(repsize, k=k)),
)


def register_image_size(D):
    def make(repsize):
        kw = _size_dependent(repsize)
        kw.update(input_width=D, input_height=D)
        return kw
    Refinement_Args["size_dependent_conv_%d" % D] = ("reg", make)
    return "size_dependent_conv_%d" % D


class RefinementNetwork(nn.Module):
    def __init__(self, input_width, input_height, input_channels, output_size, kernel_sizes, n_channels, strides,
                 paddings, hidden_sizes, lstm_size, lstm_input_size, added_fc_input_size=0, batch_norm_conv=False,
                 batch_norm_fc=False, init_w=1e-4, hidden_init=nn.init.xavier_uniform_, hidden_activation=nn.ReLU(),
                 lambda_output_activation=identity, k=None):
        assert len(kernel_sizes) == len(n_channels) == len(strides) == len(paddings)
        super().__init__()
        hidden_sizes = hidden_sizes or []
        self.hidden_sizes, self.input_width, self.input_height = hidden_sizes, input_width, input_height
        self.input_channels, self.lstm_size, self.output_size = input_channels, lstm_size, output_size
        self.lambda_output_activation, self.hidden_activation = lambda_output_activation, hidden_activation
        self.batch_norm_conv, self.batch_norm_fc, self.added_fc_input_size = batch_norm_conv, batch_norm_fc, added_fc_input_size
        self.conv_input_length = input_width * input_height * input_channels
        self.K = k
        supported = (list(kernel_sizes) == [5, 5, 5] and list(strides) == [1, 1, 1] and list(paddings) == [2, 2, 2]
                     and list(n_channels[:2]) == [32, 32] and n_channels[2] == output_size and input_channels == 17
                     and list(hidden_sizes) == [128] and lstm_size == 128 and added_fc_input_size == 0
                     and isinstance(hidden_activation, nn.ELU) and input_width == input_height
                     and lstm_input_size == output_size * 5 + 128)
        # the reference probes its conv stack with a zero image and AvgPool2d(input_width) in __init__ (:150-158): the
        # 'large_reg' (paddings 0 -> 52x52 map) and 'large_sequence' (stride 2 -> 1x1 map) registry entries fail right there
        side = input_width
        for ks, stride, pad in zip(kernel_sizes, strides, paddings):
            side = (side + 2 * pad - ks) // stride + 1
        if side < input_width:
            raise RuntimeError("Given input size: (%dx%dx%d). Calculated output size: (%dx0x0). Output size is too small "
                               "(AvgPool2d(%d) after the conv stack; this registry entry does not construct in the reference "
                               "either, refinement_network.py:157)" % (n_channels[-1], side, side, n_channels[-1], input_width))
        if not supported:
            raise NotImplementedError("only the 'size_dependent_conv' refinement architecture "
                                      "(refinement_network.py:31-46) is built")
        self.conv_layers = nn.ModuleList()
        self.conv_norm_layers, self.fc_layers, self.fc_norm_layers = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.avg_pooling = nn.AvgPool2d(kernel_size=input_width)
        self.lstm = nn.LSTM(lstm_input_size, lstm_size, num_layers=1, batch_first=True)
        cin = input_channels
        for cout, ks, stride, pad in zip(n_channels, kernel_sizes, strides, paddings):
            conv = nn.Conv2d(cin, cout, ks, stride=stride, padding=pad)
            hidden_init(conv.weight)
            conv.bias.data.fill_(0)
            self.conv_layers.append(conv)
            cin = cout
        fc_in = n_channels[-1] + added_fc_input_size          # AvgPool2d(D) leaves one value per channel
        for h in hidden_sizes:
            fc = nn.Linear(fc_in, h)
            fc.weight.data.uniform_(-init_w, init_w)
            fc.bias.data.uniform_(-init_w, init_w)
            self.fc_layers.append(fc)
            fc_in = h
        self.last_fc = nn.Linear(lstm_size, output_size)
        self.last_fc2 = nn.Linear(lstm_size, output_size)
        lin = np.linspace(-1, 1, input_width)
        self.coords = from_numpy(np.stack([np.tile(lin[None, :], (input_height, 1)),
                                           np.tile(np.linspace(-1, 1, input_height)[:, None], (1, input_width))], 0)[None])

    def forward(self, input, hidden1, hidden2, extra_input, add_fc_input=None):
        """input (N,15,D,D), hidden1/2 (N,128), extra_input (N,5*Rs) -> (lambdas1, lambdas2, h (1,N,128), c (1,N,128)).
        Inference-only standalone entry point; training goes through OP3Model.run_schedule."""
        from op3_b200.torch.op3_modules._standalone import refinement_forward
        return refinement_forward(self, input, hidden1, hidden2, extra_input)

    def initialize_hidden(self, bs):
        return ptu.zeros((1, bs, self.lstm_size)), ptu.zeros((1, bs, self.lstm_size))


class RefinementNetwork_v2_No_Sharing(nn.Module):
    """refinement_network.py:235-322: one RefinementNetwork per slot (`models.<k>...` state_dict keys); slot k refines the
    rows (b, k) and the K outputs are concatenated slot-major (output row k*bs + b), the reference's row order."""
    no_sharing = True

    def __init__(self, input_width, input_height, input_channels, output_size, kernel_sizes, n_channels, strides,
                 paddings, hidden_sizes, lstm_size, lstm_input_size, added_fc_input_size=0, batch_norm_conv=False,
                 batch_norm_fc=False, init_w=1e-4, hidden_init=nn.init.xavier_uniform_, hidden_activation=nn.ReLU(),
                 lambda_output_activation=identity, k=None):
        super().__init__()
        if k is None:
            raise ValueError("A value of k is needed to initialize this model!")
        self.K = k
        self.input_width, self.lstm_size = input_width, lstm_size
        self.models = nn.ModuleList([
            RefinementNetwork(input_width, input_height, input_channels, output_size, kernel_sizes, n_channels, strides,
                              paddings, hidden_sizes, lstm_size, lstm_input_size, added_fc_input_size, batch_norm_conv,
                              batch_norm_fc, init_w, hidden_init, hidden_activation, lambda_output_activation, k=None)
            for _ in range(k)])

    def forward(self, input, hidden1, hidden2, extra_input=None, add_fc_input=None):
        """input (B*K,15,D,D), hidden1/2 (B*K,128), extra_input (B*K,5*Rs) -> (lambdas1, lambdas2, h, c), rows slot-major."""
        g = self._get_ith_input
        c = lambda t: None if t is None else t.contiguous()
        outs = [self.models[i](c(g(input, i)), c(g(hidden1, i)), c(g(hidden2, i)), c(g(extra_input, i)), c(g(add_fc_input, i)))
                for i in range(self.K)]
        # the per-slot hidden outputs are (1, bs, lstm): torch.cat gives (K, bs, lstm) as in the reference (:304-305), which
        # OP3Model.refine flattens to rows k*bs + b (op3_model.py:240-241)
        return (torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs]),
                torch.cat([o[2] for o in outs]), torch.cat([o[3] for o in outs]))

    def _get_ith_input(self, x, i):
        return None if x is None else x.view([-1, self.K] + list(x.shape[1:]))[:, i]

    def initialize_hidden(self, bs):
        return ptu.zeros((1, bs, self.lstm_size)), ptu.zeros((1, bs, self.lstm_size))
