"""CPU, world_size 2 over gloo: host-side logic of the multi-GPU path (gradient mean, candidate sharding / gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from op3_b200 import parallel as par
    # shard_range partitions exactly
    n = 11
    s, e = par.shard_range(n)
    assert (s, e) == ((0, 6) if rank == 0 else (6, 11))
    # gradient mean: per-shard gradients (local-batch normalised, like a DataParallel replica) -> mean over ranks
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    par.allreduce_mean_(g)
    assert torch.allclose(g, torch.arange(10, dtype=torch.float32) * 1.5)
    # uneven gather of per-candidate costs: every rank ends up with the same full vector and ranking
    full = torch.tensor([5., 1., 9., 3., 7., 2., 8., 0., 6., 4., 10.])
    got = par.all_gather_rows(full[s:e].clone(), n)
    assert torch.equal(got, full)
    order = got.sort()[1]
    assert order[0].item() == 7
    gathered2d = par.all_gather_rows(torch.full((e - s, 3), float(rank)), n)
    assert gathered2d.shape == (n, 3) and gathered2d[:6].eq(0).all() and gathered2d[6:].eq(1).all()
    # device-resident data feed: the sampler gives every rank a disjoint shard of the same global batch (host logic
    # only: uint8 tensors on the CPU here; the conversion kernel is CUDA-only and covered by the GPU tests)
    from op3_b200.torch.data_management.dataset import DeviceBlocksDataset
    frames = torch.arange(12, dtype=torch.uint8).view(12, 1, 1, 1, 1).expand(12, 2, 3, 4, 4).contiguous()
    acts = torch.arange(12, dtype=torch.float32).view(12, 1, 1).expand(12, 2, 4).contiguous()
    ds = DeviceBlocksDataset(frames, acts, batchsize=3, shuffle=True, seed=5)
    assert ds.global_batch == 6 and len(ds.dataloader) == 2 and ds.action_dim == 4
    seen = []
    for fr, ac in ds.dataloader:
        assert fr.shape == (3, 2, 3, 4, 4) and fr.dtype == torch.uint8 and ac.shape == (3, 2, 4)
        assert torch.equal(fr[:, 0, 0, 0, 0].float(), ac[:, 0, 0])
        seen += fr[:, 0, 0, 0, 0].tolist()
    allseen = [None, None]
    dist.all_gather_object(allseen, seen)
    assert sorted(allseen[0] + allseen[1]) == list(range(12)), allseen
    assert not set(allseen[0]) & set(allseen[1])
    out[rank] = 1
    dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_shard_range_covers_everything():
    from op3_b200.parallel import shard_range
    for n in (0, 1, 7, 4096, 4099):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
