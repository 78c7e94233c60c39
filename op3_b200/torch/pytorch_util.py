"""Device shims mirroring op3/torch/pytorch_util.py:70-144 of the reference (same names and meaning).

A module-global `device` (None until set_gpu_mode) decides where the tensor factories allocate.  The OP3 kernels
are CUDA-only, so unlike the reference a model built while `device` is None must be moved to a GPU before use."""
import numpy as np
import torch

_use_gpu = False
device = None
_gpu_id = 0


def set_gpu_mode(mode, gpu_id=0):
    global _use_gpu, device, _gpu_id
    _gpu_id, _use_gpu = gpu_id, mode
    device = torch.device("cuda:%d" % gpu_id if mode else "cpu")


def gpu_enabled():
    return _use_gpu


def set_device(gpu_id):
    torch.cuda.set_device(gpu_id)


def _dev(torch_device):
    return device if torch_device is None else torch_device


def fanin_init(tensor):
    """U(-1/sqrt(fan_in), 1/sqrt(fan_in)) with the reference's fan-in convention (pytorch_util.py:42-51:
    size[0] for 2-D weights, prod(size[1:]) otherwise)."""
    size = tensor.size()
    if len(size) == 2:
        fan_in = size[0]
    elif len(size) > 2:
        fan_in = int(np.prod(size[1:]))
    else:
        raise Exception("Shape must be have dimension at least 2.")
    bound = 1.0 / np.sqrt(fan_in)
    return tensor.data.uniform_(-bound, bound)


def FloatTensor(*args, torch_device=None, **kwargs):
    return torch.FloatTensor(*args, **kwargs, device=_dev(torch_device))


def from_numpy(*args, **kwargs):
    return torch.from_numpy(*args, **kwargs).float().to(device)


def get_numpy(tensor):
    return tensor.to("cpu").detach().numpy()


def zeros(*sizes, torch_device=None, **kwargs):
    return torch.zeros(*sizes, **kwargs, device=_dev(torch_device))


def ones(*sizes, torch_device=None, **kwargs):
    return torch.ones(*sizes, **kwargs, device=_dev(torch_device))


def ones_like(*args, torch_device=None, **kwargs):
    return torch.ones_like(*args, **kwargs, device=_dev(torch_device))


def zeros_like(*args, torch_device=None, **kwargs):
    return torch.zeros_like(*args, **kwargs, device=_dev(torch_device))


def randn(*args, torch_device=None, **kwargs):
    return torch.randn(*args, **kwargs, device=_dev(torch_device))


def tensor(*args, torch_device=None, **kwargs):
    return torch.tensor(*args, **kwargs, device=_dev(torch_device))


def normal(*args, **kwargs):
    return torch.normal(*args, **kwargs).to(device)


def check_nan(list_of_tensors, folder=None):
    """Raises instead of dropping into pdb (pytorch_util.py:6-17)."""
    for t in list_of_tensors:
        if t is not None and torch.isnan(t).any():
            if folder:
                open(folder + "/NAN_ALERT.txt", "w").close()
            raise FloatingPointError("NaN detected")
