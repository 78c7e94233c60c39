"""BroadcastCNN parameter container mirroring op3/torch/conv_networks.py:263-350 (the only class of that file OP3
uses).  The spatial broadcast + coordinate concat + conv stack run in op3_decoder_fwd/bwd (csrc/decoder.cu,
csrc/conv5x5.cu); layer 1 is algebraically collapsed so the (N,130,D,D) broadcast input is never materialised."""
import numpy as np
from torch import nn as nn

from op3_b200.torch.networks import identity
from op3_b200.torch.pytorch_util import from_numpy


class BroadcastCNN(nn.Module):
    def __init__(self, input_width, input_height, input_channels, output_size, kernel_sizes, n_channels, strides,
                 paddings, hidden_sizes=None, added_fc_input_size=0, batch_norm_conv=False, batch_norm_fc=False,
                 init_w=1e-4, hidden_init=nn.init.xavier_uniform_, hidden_activation=nn.ReLU(),
                 output_activation=identity):
        assert len(kernel_sizes) == len(n_channels) == len(strides) == len(paddings)
        super().__init__()
        self.hidden_sizes = hidden_sizes or []
        self.input_width, self.input_height, self.input_channels = input_width, input_height, input_channels
        self.output_size, self.output_activation, self.hidden_activation = output_size, output_activation, hidden_activation
        self.batch_norm_conv, self.batch_norm_fc, self.added_fc_input_size = batch_norm_conv, batch_norm_fc, added_fc_input_size
        self.conv_input_length = input_width * input_height * input_channels
        self.conv_layers = nn.ModuleList()
        self.conv_norm_layers, self.fc_layers, self.fc_norm_layers = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        cin = input_channels
        for cout, ks, stride, pad in zip(n_channels, kernel_sizes, strides, paddings):
            conv = nn.Conv2d(cin, cout, ks, stride=stride, padding=pad)
            hidden_init(conv.weight)
            conv.bias.data.fill_(0)
            self.conv_layers.append(conv)
            cin = cout
        lin_x = np.linspace(-1, 1, input_width)
        lin_y = np.linspace(-1, 1, input_height)
        xs = np.tile(lin_x[None, :], (input_height, 1))
        ys = np.tile(lin_y[:, None], (1, input_width))
        self.coords = from_numpy(np.stack([xs, ys], 0)[None])       # plain attribute like the reference (:324)

    def check_supported(self):
        ks = [c.kernel_size for c in self.conv_layers]
        ok = (len(self.conv_layers) == 5 and all(k == (5, 5) for k in ks)
              and [c.out_channels for c in self.conv_layers] == [32, 32, 32, 32, 4]
              and all(c.stride == (1, 1) and c.padding == (2, 2) for c in self.conv_layers)
              and isinstance(self.hidden_activation, nn.ELU) and self.input_width == self.input_height)
        if not ok:
            raise NotImplementedError("only the 'reg' decoder architecture (5x 5x5/s1/p2 convs 32,32,32,32,4 + ELU, "
                                      "decoder_network.py:11-23) is built")
