"""CPU: the drop-in boundary without a GPU -- shipped checkpoints load strictly into the mirror modules, and the
reference's OWN training script (exps/train_op3/train_op3.py, unmodified, read from /root/reference when that checkout is
present) runs on top of `python -m op3_b200.run` through argument parsing, the launcher, `load_dataset` and
`create_model_v2` up to the first CUDA call (`m.cuda()`, train_op3.py:101), which cannot succeed in a GPU-less container."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from helpers import load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
VARIANT = dict(refinement_model_type="size_dependent_conv", decoder_model_type="reg", dynamics_model_type="reg_ac32",
               sto_repsize=64, det_repsize=64, extra_args=dict(beta=1e-2, deterministic_sampling=False))


def test_shipped_checkpoints_load_strict():
    """The three shipped checkpoints (weights committed in tests/golden/pretrained_*.npz): 73 / 85 / 85 tensors,
    load_state_dict(strict=True) into the mirror model, values preserved."""
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    for name, n in [("stack", 73), ("pickplace", 85), ("cloth", 85)]:
        g = load_golden("pretrained_" + name)
        sd = {k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w/")}
        assert len(sd) == n
        m = create_model_v2(dict(VARIANT, K=int(g["K"])), 64, 64, int(g["action_dim"]))
        m.load_state_dict(sd, strict=True)
        for k, v in m.state_dict().items():
            assert torch.equal(v, sd[k]), k


def test_model_pickles_without_engine_and_refuses_multi_gpu_dataparallel(monkeypatch):
    import copy
    import pickle
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    m = create_model_v2(dict(VARIANT, K=3), 64, 64, 0)
    m._engine = object()                                    # stands for a live Engine (ctypes handle + device buffers)
    m2 = pickle.loads(pickle.dumps(m))
    assert m2._engine is None and copy.deepcopy(m)._engine is None
    assert torch.equal(m2.initial_lambdas1, m.initial_lambdas1)
    monkeypatch.setattr(torch.cuda, "device_count", lambda: 2)
    with pytest.raises(RuntimeError, match="one process per GPU"):
        m._replicate_for_data_parallel()


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "exps", "train_op3")), reason="reference checkout not present")
def test_reference_train_script_runs_unchanged_until_cuda(tmp_path):
    driver = tmp_path / "drive.py"
    driver.write_text(textwrap.dedent('''
        import sys, types
        import numpy as np
        import torch

        # --- environment stand-ins for packages the reference needs but this image lacks (not reference code)
        h5 = types.ModuleType("h5py")
        class File(dict):
            def __init__(self, path, mode="r"):
                rng = np.random.RandomState(0)
                T, A = 6, (6 if "pickplace" in path else 4)
                split = lambda n: {"features": rng.randint(0, 256, (T, n, 64, 64, 3)).astype(np.uint8),
                                   "actions": rng.randn(T, n, A).astype(np.float32)}
                super().__init__(training=split(24), validation=split(8))
        h5.File = File
        sys.modules["h5py"] = h5
        sys.modules["git"] = None            # launcher_util.py:543-569: no gitpython => git_infos = None

        sys.path.insert(0, %(ref)r)
        import op3.launchers.conf as conf
        conf.LOCAL_LOG_DIR = %(log)r         # the reference checkout is read-only here

        class ReachedCuda(Exception):
            pass
        def fake_cuda(self, *a, **k):
            raise ReachedCuda(type(self).__module__ + "." + type(self).__name__)
        torch.nn.Module.cuda = fake_cuda

        import op3_b200.run
        try:
            op3_b200.run.main([%(script)r, "-va", "pickplace", "-de", "1"])
        except (ReachedCuda, RuntimeError) as e:
            # the launcher calls ptu.set_gpu_mode(True) (launcher_util.py:150), so the first CUDA touch is either a
            # tensor created on ptu.device inside OUR create_model_v2 or train_vae's m.cuda()
            import traceback
            tb = "".join(traceback.format_exception(type(e), e, e.__traceback__))
            if not isinstance(e, ReachedCuda) and "NVIDIA driver" not in str(e) and "CUDA" not in str(e):
                raise
            import op3.torch.op3_modules.op3_model as aliased
            import op3_b200.torch.op3_modules.op3_model as ours
            assert aliased is ours
            assert "train_op3.py" in tb and "train_vae" in tb
            print("REACHED_CUDA", "op3_b200/torch/op3_modules/op3_model.py" in tb or str(e))
    ''') % dict(ref=REF, log=str(tmp_path / "logs"), script=os.path.join(REF, "exps", "train_op3", "train_op3.py")))
    env = dict(os.environ, PYTHONPATH=ROOT, PYTHONDONTWRITEBYTECODE="1", PYTHONWARNINGS="ignore")
    r = subprocess.run([sys.executable, str(driver)], capture_output=True, text=True, timeout=600, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    # train_vae got through load_dataset (our BlocksDataset), create_model_v2 (our OP3Model) and nn.DataParallel
    assert "REACHED_CUDA" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
