"""CPU: the C-ABI library loads and exports every symbol include/op3_b200.h declares (no compute calls), the Python
binding declares the same set, and the host-side scheduler logic mirrors the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "op3_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(op3_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from op3_b200 import _build, _lib
    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "libop3_b200.so does not export %s" % name
    assert sorted(_lib.SIGNATURES) == declared, "ctypes SIGNATURES and the header disagree"
    bound = _lib.load()
    assert b"sm_100a" in bound.op3_version()


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    from op3_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU"):
        _lib.load()


def test_size_queries_do_not_need_a_gpu():
    from op3_b200 import _lib
    lib = _lib.load()
    d = _lib.Op3Dims(B=8, K=7, D=64, Rd=64, Rs=64, A=0, beta=0.01)
    N = 56
    # T [N][800] + a1..a4 + 8 slots of per-image operand maxima (fp16 conv mode)
    assert lib.op3_decoder_acts_floats(ctypes.byref(d), N) == N * (800 + 4 * 32 * 64 * 64) + 8 * ((N + 3) // 4 * 4)
    assert lib.op3_prepared_floats(ctypes.byref(d)) > 25 * 32 * 128
    bad = _lib.Op3Dims(B=8, K=7, D=60, Rd=64, Rs=64, A=0, beta=0.01)
    assert lib.op3_prepared_floats(ctypes.byref(bad)) == 0
    assert b"multiple of 8" in lib.op3_last_error()


def test_state_dict_keys_match_reference_layout():
    """73 / 85 entries with the reference key names, incl. the `inital_deter_state` typo."""
    from oracle import op3_oracle as orc
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    for action_dim, n in ((0, 73), (4, 85)):
        v = dict(refinement_model_type="size_dependent_conv", decoder_model_type="reg", dynamics_model_type="reg_ac32",
                 K=4, extra_args=dict(beta=1e-2, deterministic_sampling=False))
        m = create_model_v2(v, 64, 64, action_dim)
        sd = m.state_dict()
        shapes = dict(orc.param_shapes(action_dim))
        assert len(sd) == n and set(sd) == set(shapes)
        for k, t in sd.items():
            assert tuple(t.shape) == tuple(shapes[k]), k
        assert "inital_deter_state" in sd
        assert sum(p.numel() for p in m.parameters()) == (876305 if action_dim == 0 else 938930)


def test_training_scheduler_matches_reference_fixture(golden_dir):
    from op3_b200.torch.op3_modules.op3_trainer import TrainingScheduler
    fx = np.load(os.path.join(golden_dir, "schedules.npz"), allow_pickle=True)
    cases = fx["cases"]
    for kind, epoch, is_train, max_T, expect, save in cases:
        np.random.seed(3)
        s = TrainingScheduler(4, kind, max_T, 5)
        got = s.get_schedule(int(epoch), bool(is_train))
        assert np.array_equal(got, expect), (kind, epoch, is_train, max_T)
        assert s.should_save_model(int(epoch)) == save
        assert np.array_equal(s.get_loss_schedule(got), np.arange(1, len(got) + 1))


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU oracle port timed on the host cores) must print ONE JSON line with the keys the
    driver reads, without touching a GPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "train_sequences_per_sec" and j["unit"] == "sequences/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["steps"] == 1
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]


def test_install_aliases_and_device_dataset_host_logic():
    """`op3_b200.install()` makes the reference's module names resolve to this package, including the data
    feed; the device-resident dataset's sampler covers every item once per epoch (single process, CPU tensors)."""
    import importlib
    import sys
    import torch
    import op3_b200
    saved = {k: sys.modules.get(k) for k in op3_b200._ALIASES}
    try:
        op3_b200.install()
        for ref_name, ours in op3_b200._ALIASES.items():
            assert sys.modules[ref_name] is importlib.import_module(ours), ref_name
        from op3.torch.data_management.dataset import BlocksDataset, DeviceBlocksDataset   # noqa: F401 (resolves through the alias)
        from op3.torch.op3_modules.op3_trainer import TrainingScheduler
        assert list(TrainingScheduler(4, "curriculum", 5, 5).get_schedule(0, is_train=False)) == [0, 0, 0, 0, 1, 0]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    from op3_b200.torch.data_management.dataset import DeviceBlocksDataset
    frames = torch.arange(10, dtype=torch.uint8).view(1, 10, 1, 1, 1).expand(3, 10, 4, 4, 3).contiguous()   # HDF5 layout (T,N,D,D,3)
    ds = DeviceBlocksDataset(frames, None, batchsize=4, shuffle=True, seed=1, drop_last=False, time_major=True)
    assert ds.n == 10 and ds.action_dim == 0 and len(ds.dataloader) == 3
    seen = []
    for (fr,) in ds.dataloader:
        assert fr.shape[1:] == (3, 4, 4, 3) and fr.dtype == torch.uint8
        seen += fr[:, 0, 0, 0, 0].tolist()
    assert sorted(seen) == list(range(10))
    first = [b[0][:, 0, 0, 0, 0].tolist() for b in ds.dataloader]
    ds2 = DeviceBlocksDataset(frames, None, batchsize=4, shuffle=True, seed=1, drop_last=False, time_major=True)
    ds2.epoch = 1
    assert [b[0][:, 0, 0, 0, 0].tolist() for b in ds2.dataloader] == first     # the permutation is a function of (seed, epoch)
