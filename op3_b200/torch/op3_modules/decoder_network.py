"""Mirror of op3/torch/op3_modules/decoder_network.py: registry `Decoder_Args` (:10-38) and DecoderNetwork (:42-88).

DecoderNetwork(latents (N,R)) -> (mask_logits (N,1,D,D), colors (N,3,D,D)) runs op3_decoder_fwd (csrc/decoder.cu);
inside OP3Model the decoder is driven by the schedule engine with its gradient kernels."""
import ctypes as C

import torch
from torch import nn

from op3_b200 import _lib
from op3_b200.torch.networks import identity
from op3_b200.torch.conv_networks import BroadcastCNN


def _reg(full_rep_size, **extra):
    return dict(hidden_sizes=[], output_size=64 * 64 * 3, input_width=64, input_height=64,
                input_channels=full_rep_size + 2, kernel_sizes=[5] * 5, n_channels=[32, 32, 32, 32, 4],
                strides=[1] * 5, paddings=[2] * 5, hidden_activation=nn.ELU(), **extra)


# name -> (class kind, lambda -> constructor kwargs); the extension point for other image sizes
Decoder_Args = dict(
    reg=("reg", lambda full_rep_size: _reg(full_rep_size)),
    reg_no_share=("reg_no_share", lambda full_rep_size, k: _reg(full_rep_size, k=k)),
)


def register_image_size(D):
    """Add `reg_<D>` (e.g. 128x128 frames) to the registry, as the reference would by editing its dict."""
    def make(full_rep_size):
        kw = _reg(full_rep_size)
        kw.update(input_width=D, input_height=D, output_size=D * D * 3)
        return kw
    Decoder_Args["reg_%d" % D] = ("reg", make)
    return "reg_%d" % D


class DecoderNetwork(nn.Module):
    def __init__(self, input_width, input_height, input_channels, output_size, kernel_sizes, n_channels, strides,
                 paddings, hidden_sizes=None, added_fc_input_size=0, batch_norm_conv=False, batch_norm_fc=False,
                 init_w=1e-4, hidden_init=nn.init.xavier_uniform_, hidden_activation=nn.ReLU(),
                 output_activation=identity):
        super().__init__()
        self.broadcast_net = BroadcastCNN(input_width, input_height, input_channels, output_size, kernel_sizes,
                                          n_channels, strides, paddings, hidden_sizes, added_fc_input_size,
                                          batch_norm_conv, batch_norm_fc, init_w, hidden_init, hidden_activation,
                                          output_activation)
        self.broadcast_net.check_supported()
        self.input_width = input_width

    def forward(self, latents):
        """latents (N,R) -> mask_logits (N,1,D,D), colors (N,3,D,D).  Inference-only entry point (no autograd graph);
        training goes through OP3Model.run_schedule, which owns the decoder backward kernels."""
        from op3_b200.torch.op3_modules._standalone import decoder_forward
        return decoder_forward(self, latents)


def _ith_slot(x, K, i):
    """rows (b, i) of a (bs*K, ...) tensor (decoder_network.py:156-162)."""
    return None if x is None else x.view([-1, K] + list(x.shape[1:]))[:, i]


class DecoderNetwork_v2_No_Sharing(nn.Module):
    """decoder_network.py:92-162: one DecoderNetwork per slot (`models.<k>.broadcast_net...` state_dict keys).  Slot k decodes
    the rows (b, k) of the (b, k)-major input; the K outputs are concatenated slot-major, i.e. output row k*bs + b holds
    Dec_k(latents[b*K + k]) -- the reference's row order, which OP3Model then reads as (b, k)-major again (kept as is).
    Inside OP3Model the engine runs the same kernels once per slot with that slot's weights (engine.py)."""
    no_sharing = True

    def __init__(self, input_width, input_height, input_channels, output_size, kernel_sizes, n_channels, strides,
                 paddings, hidden_sizes=None, added_fc_input_size=0, batch_norm_conv=False, batch_norm_fc=False,
                 init_w=1e-4, hidden_init=nn.init.xavier_uniform_, hidden_activation=nn.ReLU(),
                 output_activation=identity, k=None):
        super().__init__()
        if k is None:
            raise ValueError("A value of k is needed to initialize this model!")
        self.K = k
        self.input_width = input_width
        self.models = nn.ModuleList([
            DecoderNetwork(input_width, input_height, input_channels, output_size, kernel_sizes, n_channels, strides, paddings,
                           hidden_sizes, added_fc_input_size, batch_norm_conv, batch_norm_fc, init_w, hidden_init,
                           hidden_activation, output_activation) for _ in range(k)])

    def forward(self, latents):
        """latents (B*K,R) -> mask_logits (B*K,1,D,D), colors (B*K,3,D,D), rows slot-major (see class docstring)."""
        outs = [self.models[i](self._get_ith_input(latents, i).contiguous()) for i in range(self.K)]
        return torch.cat([o[0] for o in outs]), torch.cat([o[1] for o in outs])

    def _get_ith_input(self, x, i):
        return _ith_slot(x, self.K, i)
