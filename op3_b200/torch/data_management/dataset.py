"""Mirror of op3/torch/data_management/dataset.py (Dataset :9-31, BlocksDataset :34-50) plus a device-resident feed.

The reference materialises the whole HDF5 split as float32 `TensorDataset`s on the host (train_op3.py:17-75: 4 bytes
per sub-pixel, `pin_memory` DataLoader, `.to(device) / 255` per batch).  At > 1000 sequences/s per GPU that feed is the
bottleneck, so `DeviceBlocksDataset` keeps the frames as uint8 ON THE GPU in the file's own layout (T, N, D, D, 3) --
10 000 pickplace sequences of 21 frames are 2.6 GB of the 180 GB -- draws batches with a per-rank sharded sampler and
converts each batch with one kernel (op3_prepare_frames: /255 + channels-last -> planar).  Reading the HDF5 file
itself stays with the reference's loader (h5py); `from_arrays` takes what `np.array(hdf5_file[split]['features'])`
returns."""
import numpy as np
import torch
from torch.utils.data.dataloader import DataLoader
from torch.utils.data.dataset import ConcatDataset

from op3_b200 import parallel as par
from op3_b200.torch.op3_modules.op3_trainer import frames_to_float


class Dataset:
    def __init__(self, torch_dataset, batchsize=8, shuffle=True):
        self.dataset = torch_dataset
        self.batchsize = batchsize
        self.shuffle = shuffle
        self._dataloader = DataLoader(torch_dataset, batch_size=batchsize, shuffle=shuffle, pin_memory=True)

    def set_batchsize(self, batchsize):
        self._dataloader = DataLoader(self.dataset, batch_size=batchsize, shuffle=self.shuffle)

    def add(self, dataset):
        self.dataset = ConcatDataset([self.dataset, dataset])
        self._dataloader = DataLoader(self.dataset, batch_size=self.batchsize, shuffle=self.shuffle)

    @property
    def dataloader(self):
        return self._dataloader


class BlocksDataset(Dataset):
    def __init__(self, torch_dataset, batchsize=8, shuffle=True):
        super().__init__(torch_dataset, batchsize, shuffle)
        self.action_dim = self.dataset.tensors[1].shape[-1] if len(self.dataset.tensors) == 2 else 0

    def __getitem__(self, idx):
        frames = self._dataloader.dataset[idx][0]            # (T, 3, D, D)
        if self.action_dim == 0:
            return frames, None
        return frames, self._dataloader.dataset[idx][1]      # (T-1, A) or (T, A)


class _DeviceLoader:
    """Iterable over the batches of one epoch: a seeded permutation shared by all ranks, of which every rank takes its
    contiguous shard of each global batch (equal local batches, like DataParallel's scatter)."""

    def __init__(self, ds):
        self.ds = ds

    def __len__(self):
        return self.ds.n // self.ds.global_batch if self.ds.drop_last else -(-self.ds.n // self.ds.global_batch)

    def __iter__(self):
        ds = self.ds
        if ds.shuffle:
            g = torch.Generator().manual_seed(ds.seed + ds.epoch)
            order = torch.randperm(ds.n, generator=g)
        else:
            order = torch.arange(ds.n)
        ds.epoch += 1
        rank, world = par.world()
        for i in range(len(self)):
            idx = order[i * ds.global_batch:(i + 1) * ds.global_batch]
            s, e = par.shard_range(idx.numel(), rank, world)
            yield ds.batch(idx[s:e])


class DeviceBlocksDataset:
    """uint8 frames resident on the GPU.  `dataloader` yields what OP3Trainer.prepare_tensors accepts -- (uint8 frames
    (b,T,D,D,3) or (b,T,3,D,D)[, actions]) -- and prepare_tensors converts the batch with one kernel."""

    def __init__(self, frames_u8, actions=None, batchsize=8, shuffle=True, seed=0, drop_last=True, time_major=False):
        """frames_u8: uint8 CUDA tensor (N,T,3,D,D), (N,T,D,D,3), or with time_major=True the HDF5 layout (T,N,D,D,3).
        actions: float CUDA tensor (N,T',A) (or (T',N,A) when time_major) or None.  batchsize is PER RANK."""
        if frames_u8.dtype != torch.uint8:
            raise ValueError("DeviceBlocksDataset wants uint8 frames (the conversion kernel runs on CUDA tensors only)")
        self.frames, self.actions, self.time_major = frames_u8, actions, time_major
        self.n = frames_u8.shape[1] if time_major else frames_u8.shape[0]
        self.batchsize, self.shuffle, self.seed, self.drop_last, self.epoch = batchsize, shuffle, seed, drop_last, 0
        self.global_batch = batchsize * par.world()[1]
        self.action_dim = actions.shape[-1] if actions is not None else 0

    @classmethod
    def from_arrays(cls, features, actions=None, device="cuda", **kw):
        """features: what the reference reads from the file, np.uint8 (T,N,D,D,3) (train_op3.py:24-30)."""
        f = torch.from_numpy(np.ascontiguousarray(features)).to(device)
        a = torch.from_numpy(np.ascontiguousarray(actions)).float().to(device) if actions is not None else None
        return cls(f, a, time_major=True, **kw)

    def batch(self, idx):
        idx = idx.to(self.frames.device)
        if self.time_major:
            fr = self.frames.index_select(1, idx).transpose(0, 1)      # (b,T,D,D,3) view; made contiguous by the converter
            ac = self.actions.index_select(1, idx).transpose(0, 1).contiguous() if self.actions is not None else None
        else:
            fr = self.frames.index_select(0, idx)
            ac = self.actions.index_select(0, idx) if self.actions is not None else None
        return (fr, ac) if ac is not None else (fr,)

    def __getitem__(self, i):
        b = self.batch(torch.tensor([i]))
        frames = frames_to_float(b[0])[0] * 255.0            # same value range as the reference dataset items
        return frames, (b[1][0] if len(b) == 2 else None)

    def set_batchsize(self, batchsize):
        self.batchsize = batchsize
        self.global_batch = batchsize * par.world()[1]

    @property
    def dataloader(self):
        return _DeviceLoader(self)
