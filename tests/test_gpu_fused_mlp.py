"""The CTA-resident MLP chains (csrc/fused_mlp.cuh: dynamics network, refinement head) against the layer-by-layer launches
of the same library (op3_set_fused_mlp(0)) and, through them, against the oracle (tests/test_gpu_kernels.py runs the default =
fused path against physics_network.py:94-141 / refinement_network.py:203-219 goldens)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import op3_oracle as orc
from gpu_util import build_model, rel_l2, max_abs, NoiseInjector

pytestmark = pytest.mark.gpu


def _dyn_both(lib, eng, B, K, action_dim, seed):
    from op3_b200._lib import check, ptr
    N, R = B * K, 128
    gen = torch.Generator().manual_seed(seed)
    z = (torch.randn(N, R, generator=gen) * 0.7).cuda()
    ac = (torch.randn(B, 4, generator=gen) * 0.5).cuda() if action_dim else None
    up = [torch.randn(N, 64, generator=gen).cuda() for _ in range(3)]
    d_oz = torch.zeros(N, R, device="cuda")
    d_oz[:, :64] = up[0]
    st = torch.cuda.current_stream().cuda_stream
    d = eng.dims(B)
    ps = eng.params_struct()
    res = []
    for fused in (0, 1):
        lib.op3_set_fused_mlp(fused)
        try:
            A = torch.zeros(lib.op3_dynamics_acts_floats(C.byref(d)), device="cuda")
            oz, o1, o2 = torch.zeros(N, R, device="cuda"), torch.empty(N, 64, device="cuda"), torch.empty(N, 64, device="cuda")
            check(lib.op3_dynamics_fwd(C.byref(d), C.byref(ps), ptr(z), ptr(ac), ptr(A), ptr(oz), ptr(o1), ptr(o2), st))
            gflat = torch.zeros_like(eng.flat_params())
            gs = eng._struct_for(gflat)
            dz = torch.empty(N, R, device="cuda")
            scratch = torch.empty(lib.op3_dynamics_scratch_floats(C.byref(d)), device="cuda")
            check(lib.op3_dynamics_bwd(C.byref(d), C.byref(ps), ptr(z), ptr(ac), ptr(A), ptr(d_oz), ptr(up[1]), ptr(up[2]),
                                       ptr(dz), C.byref(gs), ptr(scratch), st))
            probes = None
            if K > 1:
                pa = torch.empty(N, K - 1, device="cuda")
                dl = torch.empty(N, device="cuda")
                aa = torch.empty(N, device="cuda") if action_dim else None
                check(lib.op3_dynamics_probes(C.byref(d), ptr(A), ptr(aa), ptr(pa), ptr(dl), st))
                probes = (pa, dl, aa)
            torch.cuda.synchronize()
            res.append(dict(oz=oz[:, :64].clone(), o1=o1, o2=o2, dz=dz, g=gflat, acts=A, probes=probes))
        finally:
            lib.op3_set_fused_mlp(1)
    return res


@pytest.mark.parametrize("B,K,action_dim", [(5, 4, 4), (3, 7, 0), (4, 1, 4), (2, 14, 4), (3, 16, 0), (700, 4, 4), (1, 2, 4)])
def test_fused_dynamics_matches_layerwise(B, K, action_dim):
    from op3_b200 import _lib
    lib = _lib.load()
    m, p = build_model(action_dim, K, seed=11)
    ref, fus = _dyn_both(lib, m.engine, B, K, action_dim, seed=B * 31 + K)
    for k in ("oz", "o1", "o2"):
        assert max_abs(fus[k], ref[k]) <= 2e-6 * max(1.0, ref[k].abs().max().item()), k
    assert rel_l2(fus["dz"], ref["dz"]) <= 2e-6
    # the saved activations are the contract with the backward pass and with the attention probes
    assert rel_l2(fus["acts"], ref["acts"]) <= 2e-6
    if ref["probes"] is not None:
        for a, b in zip(fus["probes"], ref["probes"]):
            if a is not None:
                assert max_abs(a, b) <= 2e-6 * max(1.0, b.abs().max().item())
    views_f = dict(zip(m.engine._names, m.engine.grad_views(fus["g"])))
    views_r = dict(zip(m.engine._names, m.engine.grad_views(ref["g"])))
    tn = ref["g"].double().norm().item()
    for k, g in views_r.items():
        if not k.startswith("dynamics_net"):
            continue
        err = (views_f[k].double() - g.double()).norm().item()
        assert err <= 5e-6 * max(g.double().norm().item(), 1e-6 * tn), (k, err, g.norm().item())


def test_fused_dynamics_is_deterministic():
    from op3_b200 import _lib
    lib = _lib.load()
    m, p = build_model(4, 4, seed=12)
    a = _dyn_both(lib, m.engine, 37, 4, 4, seed=5)[1]
    b = _dyn_both(lib, m.engine, 37, 4, 4, seed=5)[1]
    assert torch.equal(a["o1"], b["o1"]) and torch.equal(a["dz"], b["dz"]) and torch.equal(a["g"], b["g"])


@pytest.mark.parametrize("action_dim,K,B,T,schedule", [(0, 3, 5, 2, [0, 0, 1, 0]), (4, 4, 3, 3, [0, 0, 1, 1, 0])])
def test_fused_heads_match_layerwise_through_a_train_step(action_dim, K, B, T, schedule):
    """Whole schedule forward + backward with the fused chains on and off: losses, masks and every parameter gradient."""
    from op3_b200 import _lib
    lib = _lib.load()
    m, p = build_model(action_dim, K, seed=13)
    frames = orc.synth_frames(B, T, seed=3).cuda()
    actions = orc.synth_actions(B, T, seed=4).cuda() if action_dim else None
    if actions is not None:
        actions = actions[..., [0, 1, 3, 4]].contiguous()
    gen = torch.Generator().manual_seed(9)
    noise = [torch.randn(B, K, 64, generator=gen) for _ in range(len(schedule) + 1)]
    lw = np.arange(1, len(schedule) + 1, dtype=np.float64)
    out = []
    for fused in (0, 1):
        lib.op3_set_fused_mlp(fused)
        try:
            m.zero_grad(set_to_none=True)
            with NoiseInjector(noise):
                res = m.run_schedule(frames, actions, None, schedule=np.array(schedule), loss_schedule=lw)
            res[3].backward()
            torch.cuda.synchronize()
            out.append((res[3].item(), res[1].detach().clone(), {k: v.grad.detach().clone() for k, v in m.named_parameters()}))
        finally:
            lib.op3_set_fused_mlp(1)
    (l0, m0, g0), (l1, m1, g1) = out
    assert abs(l1 - l0) <= 2e-6 * abs(l0)
    assert max_abs(m1, m0) <= 2e-6
    tn = torch.sqrt(sum((v.double() ** 2).sum() for v in g0.values())).item()
    for k, g in g0.items():
        err = (g1[k].double() - g.double()).norm().item()
        assert err <= 2e-5 * max(g.double().norm().item(), 1e-6 * tn), (k, err, g.norm().item())
