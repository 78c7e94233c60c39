"""Parameter containers mirroring op3/torch/modules.py (LayerNorm :19-46, LayerNorm2D :48-69).

Inside OP3Model these normalisations are fused into the mixture / refinement kernels (mix_emit_kernel,
refine_extras_kernel); the modules exist so that the state_dict keys (`gamma`, `beta`, `scale_param`,
`center_param`) and initialisers match the reference.  Their standalone forward is plain tensor arithmetic for
inspection only -- it is not on the product path."""
import torch
import torch.nn as nn


class LayerNorm(nn.Module):
    def __init__(self, features, center=True, scale=False, eps=1e-6):
        super().__init__()
        self.center, self.scale, self.eps = center, scale, eps
        self.scale_param = nn.Parameter(torch.ones(features)) if scale else None
        self.center_param = nn.Parameter(torch.zeros(features)) if center else None

    def forward(self, x):
        out = (x - x.mean(-1, keepdim=True)) / (x.std(-1, keepdim=True) + self.eps)
        if self.scale:
            out = out * self.scale_param
        if self.center:
            out = out + self.center_param
        return out


class LayerNorm2D(nn.Module):
    def __init__(self, num_features, eps=1e-5, affine=True):
        super().__init__()
        self.num_features, self.affine, self.eps = num_features, affine, eps
        if affine:
            self.gamma = nn.Parameter(torch.empty(num_features).uniform_())   # U(0,1), modules.py:56
            self.beta = nn.Parameter(torch.zeros(num_features))

    def forward(self, x):
        n = x.size(0)
        flat = x.reshape(n, -1)
        bshape = [n] + [1] * (x.dim() - 1)
        out = (x - flat.mean(1).view(bshape)) / (flat.std(1).view(bshape) + self.eps)
        if self.affine:
            cshape = [1, -1] + [1] * (x.dim() - 2)
            out = out * self.gamma.view(cshape) + self.beta.view(cshape)
        return out
