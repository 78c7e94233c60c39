"""GPU unit parity of the individual C-ABI ops against the CPU oracle (same weights, same inputs)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import op3_oracle as orc
from helpers import load_golden, check_summary
from gpu_util import build_model, rel_l2, max_abs, compare_grads

pytestmark = pytest.mark.gpu


def _lib():
    from op3_b200 import _lib
    return _lib


def test_decoder_forward_matches_oracle_and_golden():
    from op3_b200.torch.op3_modules._standalone import decoder_forward
    g = load_golden("modules")
    m, p = build_model(4, 3, seed=3)
    z = torch.from_numpy(g["dec/z"]).cuda()
    logits, colors = m.decode_net(z)
    check_summary(g, "dec/colors", colors, 2e-5, 2e-6)
    check_summary(g, "dec/logits", logits, 2e-5, 2e-6)
    ref = orc.decoder(p, torch.from_numpy(g["dec/z"]))
    assert max_abs(colors, ref[:, :3]) <= 5e-6 * ref.abs().max().item() + 1e-6
    assert max_abs(logits, ref[:, 3:4]) <= 5e-6 * ref.abs().max().item() + 1e-6
    # standalone (no OP3Model) path prepares its own derived weights
    l2, c2 = decoder_forward(m.decode_net, z, engine=None)
    assert torch.equal(l2, logits) and torch.equal(c2, colors)


def test_refinement_network_forward_matches_golden_and_oracle():
    """RefinementNetwork.forward as a standalone module (refinement_network.py:189-219) against the reference fixture."""
    g = load_golden("modules")
    m, p = build_model(4, 3, seed=3)
    gen = torch.Generator().manual_seed(int(g["ref/seed"]))
    torch.randn(6, 128, generator=gen)                      # the decoder case's draw precedes these in make_golden.py
    a = torch.randn(6, 15, 64, 64, generator=gen)
    h = torch.randn(6, 128, generator=gen) * 0.3
    c = torch.randn(6, 128, generator=gen) * 0.3
    extra = torch.randn(6, 320, generator=gen)
    l1, l2, h2, c2 = m.refinement_net(a.cuda(), h.cuda(), c.cuda(), extra.cuda())
    assert h2.shape == (1, 6, 128) and c2.shape == (1, 6, 128)
    ref = orc.refinement_net(p, a, h, c, extra)
    for name, ours, want in [("l1", l1, ref[0]), ("l2", l2, ref[1]), ("h2", h2[0], ref[2]), ("c2", c2[0], ref[3])]:
        gold = torch.from_numpy(g["ref/" + name])
        assert max_abs(want, gold) <= 1e-5, name
        assert max_abs(ours, gold) <= 2e-5 * max(1.0, gold.abs().max().item()), (name, max_abs(ours, gold))
    with pytest.raises(RuntimeError):
        m.refinement_net(a, h.cuda(), c.cuda(), extra.cuda())


def test_decoder_backward_matches_autograd():
    lib_mod = _lib()
    lib = lib_mod.load()
    from op3_b200._lib import check, ptr
    m, p = build_model(0, 2, seed=5)
    eng = m.engine
    N, D, R = 4, 64, 128
    gen = torch.Generator().manual_seed(9)
    z = (torch.randn(N, R, generator=gen) * 0.7)
    d_out = torch.randn(N, 4, D, D, generator=gen)
    q = {k: v.clone().requires_grad_(True) for k, v in p.items() if k.startswith("decode_net")}
    zr = z.clone().requires_grad_(True)
    out = orc.decoder(q, zr)
    out.backward(d_out)
    # ours
    st = torch.cuda.current_stream().cuda_stream
    d = eng.dims(2)
    prep = eng.prepared(st)
    zc = z.cuda()
    acts = torch.empty(lib.op3_decoder_acts_floats(C.byref(d), N), device="cuda")
    o = torch.empty(N, D, D, 4, device="cuda")
    check(lib.op3_decoder_fwd(C.byref(d), N, ptr(prep), ptr(zc), ptr(acts), ptr(o), st))
    assert max_abs(o.permute(0, 3, 1, 2), out) <= 1e-5
    dz = torch.zeros(N, R, device="cuda")
    scratch = torch.empty(lib.op3_decoder_scratch_floats(C.byref(d), N), device="cuda")
    pg = torch.zeros_like(prep)
    d_out_f4 = d_out.permute(0, 2, 3, 1).contiguous().cuda()
    check(lib.op3_decoder_bwd(C.byref(d), N, ptr(prep), ptr(zc), ptr(acts), ptr(d_out_f4), ptr(scratch), ptr(dz), 0, ptr(pg), st))
    gflat = torch.zeros_like(eng.flat_params())
    gs = eng._struct_for(gflat)
    check(lib.op3_finalize_grads(C.byref(d), ptr(pg), C.byref(gs), st))
    assert rel_l2(dz, zr.grad) <= 2e-5
    views = dict(zip(eng._names, eng.grad_views(gflat)))
    ref = {k: v.grad for k, v in q.items()}
    tn = torch.sqrt(sum((v.double() ** 2).sum() for v in ref.values())).item()
    report = []
    try:
        compare_grads(views, ref, tn, 5e-5, report)
    finally:
        print("\n".join(report))
    # latent-only form (prep_grad = NULL) gives the same dz and accumulates when asked to
    dz2 = torch.ones(N, R, device="cuda")
    check(lib.op3_decoder_bwd(C.byref(d), N, ptr(prep), ptr(zc), ptr(acts), ptr(d_out_f4), ptr(scratch), ptr(dz2), 1, None, st))
    assert max_abs(dz2 - 1, dz) <= 1e-6 * dz.abs().max().item() + 1e-7


def test_dynamics_forward_and_probes_match_golden():
    g = load_golden("modules")
    m, p = build_model(4, 3, seed=3)
    z = torch.from_numpy(g["dyn/z"]).cuda()
    acts = torch.from_numpy(g["dyn/actions"]).cuda()
    d, l1, l2 = m.dynamics_net(z, acts)
    for k, v in [("deter", d), ("l1", l1), ("l2", l2)]:
        assert np.abs(v.cpu().numpy() - g["dyn/" + k]).max() <= 1e-5 * max(1.0, np.abs(g["dyn/" + k]).max()), k
    sa, ia, dl = m.dynamics_net.get_all_attention_values(z, acts[:, [0, 1, 3, 4]], 3)
    for k, v in [("act_att", sa), ("pair_att", ia), ("deltas", dl)]:
        assert np.abs(v.cpu().numpy() - g["dyn/" + k]).max() <= 1e-5 * max(1.0, np.abs(g["dyn/" + k]).max()), k


def test_dynamics_backward_matches_autograd():
    lib = _lib().load()
    from op3_b200._lib import check, ptr
    for action_dim in (4, 0):
        K, B = 3, 2
        m, p = build_model(action_dim, K, seed=6)
        eng = m.engine
        N, R = B * K, 128
        gen = torch.Generator().manual_seed(4)
        z = torch.randn(N, R, generator=gen) * 0.7
        acts_b = orc.synth_actions(B, 1, seed=8)[:, 0][:, [0, 1, 3, 4]] if action_dim else None
        up = [torch.randn(N, 64, generator=gen) for _ in range(3)]
        q = {k: v.clone().requires_grad_(True) for k, v in p.items() if k.startswith("dynamics_net")}
        zr = z.clone().requires_grad_(True)
        acts_rows = None if acts_b is None else acts_b.unsqueeze(1).expand(B, K, 4).reshape(N, 4)
        deter, l1, l2 = orc.dynamics_net(q, zr, acts_rows, K)
        (deter * up[0]).sum().add((l1 * up[1]).sum()).add((l2 * up[2]).sum()).backward()
        st = torch.cuda.current_stream().cuda_stream
        d = eng.dims(B)
        ps = eng.params_struct()
        zc = z.cuda()
        ac = None if acts_b is None else acts_b.contiguous().cuda()
        A = torch.empty(lib.op3_dynamics_acts_floats(C.byref(d)), device="cuda")
        oz, o1, o2 = torch.zeros(N, R, device="cuda"), torch.empty(N, 64, device="cuda"), torch.empty(N, 64, device="cuda")
        check(lib.op3_dynamics_fwd(C.byref(d), C.byref(ps), ptr(zc), ptr(ac), ptr(A), ptr(oz), ptr(o1), ptr(o2), st))
        assert max_abs(oz[:, :64], deter) <= 1e-5 and max_abs(o1, l1) <= 1e-5 and max_abs(o2, l2) <= 1e-5
        gflat = torch.zeros_like(eng.flat_params())
        gs = eng._struct_for(gflat)
        d_oz = torch.zeros(N, R, device="cuda")
        d_oz[:, :64] = up[0].cuda()
        dz = torch.empty(N, R, device="cuda")
        scratch = torch.empty(lib.op3_dynamics_scratch_floats(C.byref(d)), device="cuda")
        u1, u2 = up[1].cuda(), up[2].cuda()      # keep alive: the call is asynchronous
        check(lib.op3_dynamics_bwd(C.byref(d), C.byref(ps), ptr(zc), ptr(ac), ptr(A), ptr(d_oz), ptr(u1), ptr(u2), ptr(dz),
                                   C.byref(gs), ptr(scratch), st))
        assert rel_l2(dz, zr.grad) <= 2e-5
        views = dict(zip(eng._names, eng.grad_views(gflat)))
        ref = {k: v.grad for k, v in q.items()}
        tn = torch.sqrt(sum((v.double() ** 2).sum() for v in ref.values())).item()
        compare_grads(views, ref, tn, 5e-5)


@pytest.mark.parametrize("fused", [False, True], ids=["two_pass", "single_launch"])
def test_mixture_forward_pieces_match_oracle(fused):
    """masks, losses, LayerNorm'd refinement inputs and the gradient seed of one decode, through the two-pass kernels
    (Op3MixtureIO.sync = NULL) and through the single-launch cluster kernel (sync word given)."""
    lib = _lib().load()
    from op3_b200._lib import check, ptr, Op3MixtureIO
    sync = torch.zeros(4, device="cuda", dtype=torch.int32)
    for K, B in ((3, 2), (4, 5), (7, 2), (11, 1)):
        m, p = build_model(0, K, seed=2)
        eng = m.engine
        N, D, Rs = B * K, 64, 64
        gen = torch.Generator().manual_seed(K)
        dec = torch.randn(N, 4, D, D, generator=gen) * 0.5
        dec[:, :3] = dec[:, :3] * 0.3 + 0.5
        img = orc.synth_frames(B, 1, seed=3)[:, 0]
        lam = [torch.randn(B, K, Rs, generator=gen) * 0.5 for _ in range(4)]
        # oracle
        decr = dec.clone().requires_grad_(True)
        dv = decr.view(B, K, 4, D, D)
        colors, logits = dv[:, :, :3], dv[:, :, 3:4]
        mask = torch.softmax(logits, 1)
        state = {"post": {"lambdas1": lam[0].clone().requires_grad_(True), "lambdas2": lam[1].clone().requires_grad_(True),
                          "samples": lam[0], "deter_state": lam[0], "extra_info": None},
                 "prior": {"lambdas1": lam[2], "lambdas2": lam[3]}}
        probs, P, kl, clog, total = orc.get_loss(state, colors, mask, img, 1e-2)
        seed_ref, = torch.autograd.grad(total, decr, retain_graph=True)
        a_ref, _, aux = orc.refine_inputs({k: v.detach() for k, v in p.items()}, state, img,
                                          (colors, mask, logits, probs, P, total), 1e-2)
        # ours
        st = torch.cuda.current_stream().cuda_stream
        d = eng.dims(B)
        ps = eng.params_struct()
        f = lambda t: t.contiguous().cuda()
        dec_f4 = f(dec.permute(0, 2, 3, 1))
        io = Op3MixtureIO()
        keep = dict(dec=dec_f4, img=f(img), l=[f(t) for t in lam], mask=torch.empty(B, K, D, D, device="cuda"),
                    colors=torch.empty(B, K, 3, D, D, device="cuda"), recon=torch.empty(B, 3, D, D, device="cuda"),
                    stats=torch.empty(lib.op3_mixture_stats_floats(C.byref(d)), device="cuda"), losses=torch.zeros(4, device="cuda"),
                    mix1=torch.empty(N, D, D, 4, device="cuda"), mix2=torch.empty(N, D, D, 4, device="cuda"),
                    smp=torch.empty(B, D, D, 4, device="cuda"), seed=torch.empty(N, D, D, 4, device="cuda"))
        io.dec_out, io.image, io.image_bstride = ptr(keep["dec"]), ptr(keep["img"]), 3 * D * D
        io.mse_image, io.mse_image_bstride = ptr(keep["img"]), 3 * D * D
        io.post_l1, io.post_l2, io.prior_l1, io.prior_l2 = [ptr(t) for t in keep["l"]]
        io.mask, io.colors, io.final_recon = ptr(keep["mask"]), ptr(keep["colors"]), ptr(keep["recon"])
        io.stats, io.losses = ptr(keep["stats"]), ptr(keep["losses"])
        io.mix1, io.mix2, io.smp, io.seed = ptr(keep["mix1"]), ptr(keep["mix2"]), ptr(keep["smp"]), ptr(keep["seed"])
        if fused:
            io.sync = ptr(sync)
            # loss-only call first (no refinement inputs): same scalars, and the sync word must come back to zero
            io2 = Op3MixtureIO.from_buffer_copy(io)
            io2.mix1 = io2.mix2 = io2.smp = io2.seed = None
            l2 = torch.zeros(4, device="cuda")
            io2.losses = ptr(l2)
            check(lib.op3_mixture_fwd(C.byref(d), C.byref(ps), C.byref(io2), 1, st))
            assert int(sync[0].item()) == 0
        check(lib.op3_mixture_fwd(C.byref(d), C.byref(ps), C.byref(io), 1, st))
        L = keep["losses"].cpu()
        if fused:
            assert int(sync[0].item()) == 0
            assert torch.equal(l2.cpu(), L), (l2, L)
        assert abs(L[0] - clog.item()) <= 2e-6 * abs(clog.item()), (L, clog)
        assert abs(L[1] - kl.item()) <= 1e-5 * abs(kl.item())
        assert abs(L[2] - total.item()) <= 2e-6 * abs(total.item())
        recon_ref = (colors * mask).sum(1)
        assert abs(L[3] - ((recon_ref - img) ** 2).mean().item()) <= 1e-5 * ((recon_ref - img) ** 2).mean().item()
        assert max_abs(keep["mask"], mask.squeeze(2)) <= 2e-6
        assert max_abs(keep["colors"], colors) == 0.0
        assert max_abs(keep["recon"], recon_ref) <= 2e-6
        # refinement-input channels: a_ref (N,15,D,D) order op3_model.py:219-229
        mix1 = keep["mix1"].permute(0, 3, 1, 2).cpu()
        mix2 = keep["mix2"].permute(0, 3, 1, 2).cpu()
        smp = keep["smp"].permute(0, 3, 1, 2).cpu()
        a = a_ref.detach()
        tol = lambda ref: 2e-4 * ref.abs().max().item() + 1e-6     # LN divides by a std: errors scale with 1/std
        assert max_abs(mix1[:, 0], a[:, 6]) <= 2e-6 and max_abs(mix1[:, 1], a[:, 8]) <= 2e-6
        assert max_abs(mix1[:, 2], a[:, 9]) <= tol(a[:, 9]), "LN(mask_grad)"
        assert max_abs(mix1[:, 3], a[:, 11]) <= tol(a[:, 11]), "LN(leave_out)"
        assert max_abs(mix2[:, :3], a[:, 12:15]) <= tol(a[:, 12:15]), "LN(x_hat_grad)"
        assert max_abs(smp[:, :3], img) == 0.0
        assert max_abs(smp[:, 3], a.view(B, K, 15, D, D)[:, 0, 10]) <= tol(a[:, 10]), "LN(P)"
        seed = keep["seed"].permute(0, 3, 1, 2).cpu()
        assert rel_l2(seed, seed_ref) <= 2e-5


def test_mpc_cost_kernel_matches_oracle_ranking():
    from op3_b200.mpc import Cost
    gen = torch.Generator().manual_seed(3)
    Na, K, G = 64, 4, 4
    pred = torch.rand(Na, K, 3, 64, 64, generator=gen)
    goal = torch.rand(G, 3, 64, 64, generator=gen)
    final = torch.rand(Na, 3, 64, 64, generator=gen)
    gimg = torch.rand(1, 3, 64, 64, generator=gen)
    for core, agg in [("final_recon", "sum"), ("subimage", "min"), ("subimage", "sum")]:
        sc, order, gidx = orc.mpc_cost(goal, gimg, pred, final, core, agg)
        c = Cost(None, core_type=core, aggregate=agg)
        s2, o2, g2 = c.get_action_rankings(None, goal.cuda(), gimg.cuda(), None, pred.cuda(), final.cuda(), plot_actions=0)
        assert np.array_equal(o2, order.numpy()), (core, agg)
        assert np.abs(s2 - sc.numpy()).max() <= 1e-6
        if agg == "min":
            assert np.array_equal(g2, gidx.numpy())
        # bit-reproducible run to run
        s3, o3, _ = c.get_action_rankings(None, goal.cuda(), gimg.cuda(), None, pred.cuda(), final.cuda(), plot_actions=0)
        assert np.array_equal(s2, s3) and np.array_equal(o2, o3)


def test_cem_refit_and_resample_match_numpy():
    """CEM refit (mpc_pickplace.py:376-381: mean, population std + 0.05 of the best 10 %) and the clipped Gaussian
    resampling (block_pick_and_place.py:390-399) against the reference's numpy formulas on the same numbers."""
    from op3_b200.mpc import cem_refit, cem_resample
    rng = np.random.RandomState(0)
    for Na, T, A, min_std in ((4096, 1, 4, 0.05), (500, 3, 4, 0.05), (30, 1, 6, 0.0)):
        acts = rng.uniform(-2.5, 4.0, size=(Na, T, A)).astype(np.float32)
        costs = rng.rand(Na).astype(np.float32)
        order = np.argsort(costs, kind="stable")
        F = max(int(Na * 0.1), 1)
        best = acts[order][:F].astype(np.float64)
        mean_ref, std_ref = best.mean(0), best.std(0) + min_std
        mean, std = cem_refit(torch.from_numpy(acts).cuda(), torch.from_numpy(order).cuda(), F, min_std)
        assert mean.shape == (T, A) and std.shape == (T, A)
        assert np.allclose(mean.cpu().numpy(), mean_ref, rtol=1e-6, atol=1e-7)
        assert np.allclose(std.cpu().numpy(), std_ref, rtol=1e-6, atol=1e-7)
        noise = rng.randn(Na, T, A).astype(np.float32)
        lo = np.array([-2.5, 1.0, -2.5, 1.0] + [0.0] * (A - 4), dtype=np.float32)
        hi = np.array([2.5, 4.0, 2.5, 4.0] + [9.0] * (A - 4), dtype=np.float32)
        out = cem_resample(mean, std, torch.from_numpy(noise).cuda(), torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda())
        ref = np.clip(mean.cpu().numpy()[None] + std.cpu().numpy()[None] * noise, lo, hi)
        assert np.allclose(out.cpu().numpy(), ref, rtol=1e-6, atol=1e-6)
        assert out.min().item() >= -2.5 and out.max().item() <= 9.0
        # bit-reproducible
        mean2, std2 = cem_refit(torch.from_numpy(acts).cuda(), torch.from_numpy(order).cuda(), F, min_std)
        assert torch.equal(mean, mean2) and torch.equal(std, std2)
    with pytest.raises(ValueError):
        cem_refit(torch.zeros(8, 1, 4, device="cuda"), torch.arange(8, device="cuda"), 9)


def test_prepare_frames_and_device_dataset():
    """Data feed: uint8 -> float/255 in both layouts, bit-exact against torch's division,
    and the device-resident dataset handing batches to OP3Trainer.prepare_tensors."""
    from op3_b200.torch.op3_modules.op3_trainer import frames_to_float, OP3Trainer, TrainingScheduler
    from op3_b200.torch.data_management.dataset import DeviceBlocksDataset
    g = torch.Generator().manual_seed(0)
    T, N, D = 3, 10, 64
    feats = torch.randint(0, 256, (T, N, D, D, 3), generator=g, dtype=torch.uint8)      # the HDF5 layout
    ref = feats.permute(1, 0, 4, 2, 3).float() / 255.0                                  # train_op3.py:27-28 + /255
    planar = feats.permute(1, 0, 4, 2, 3).contiguous().cuda()
    assert torch.equal(frames_to_float(planar).cpu(), ref)
    cl = feats.permute(1, 0, 2, 3, 4).contiguous().cuda()
    assert torch.equal(frames_to_float(cl).cpu(), ref)
    acts = torch.rand(T, N, 4, generator=g)
    ds = DeviceBlocksDataset.from_arrays(feats.numpy(), acts.numpy(), batchsize=4, shuffle=False)
    m, _ = build_model(4, 4, seed=0)
    tr = OP3Trainer(ds, ds, m, TrainingScheduler(4, "curriculum", T, T), 4, lr=3e-4)
    batches = list(ds.dataloader)
    assert len(batches) == 2
    imgs, a = tr.prepare_tensors(batches[1])
    assert imgs.shape == (4, T, 3, D, D) and torch.equal(imgs.cpu(), ref[4:8]) and torch.equal(a.cpu(), acts.permute(1, 0, 2)[4:8])
    item, ai = ds[7]
    assert torch.allclose(item.cpu(), ref[7] * 255.0) and torch.equal(ai.cpu(), acts[:, 7])
    with pytest.raises(RuntimeError):
        frames_to_float(planar.cpu())


def test_quint_pair_kernel_is_bit_identical_to_single_cta():
    """The cta_group::2 form of the quint-tap kernel (CTA pairs sharing the weight operand) issues the same MMAs in the same
    order on the same operands as the single-CTA form: decoder forward + latent gradient must agree bit for bit."""
    import op3_b200
    lib = _lib().load()
    from op3_b200._lib import check, ptr
    op3_b200.set_precision("f16x3")
    try:
        for K, B in ((4, 6), (3, 2), (7, 9)):
            m, p = build_model(0, K, seed=9)
            eng = m.engine
            N, R = B * K, 128
            z = (torch.randn(N, R, generator=torch.Generator().manual_seed(K)) * 0.7).cuda()
            st = torch.cuda.current_stream().cuda_stream
            d = eng.dims(B)
            prep = eng.prepared(st)
            seed = torch.randn(N, 64 * 64, 4, generator=torch.Generator().manual_seed(B)).cuda()
            res = []
            for pair in (0, 1):
                lib.op3_set_tcq_pair(pair)
                acts = torch.empty(lib.op3_decoder_acts_floats(C.byref(d), N), device="cuda")
                out = torch.empty(N, 64 * 64, 4, device="cuda")
                check(lib.op3_decoder_fwd(C.byref(d), N, ptr(prep), ptr(z), ptr(acts), ptr(out), st))
                scratch = torch.empty(lib.op3_decoder_scratch_floats(C.byref(d), N), device="cuda")
                dz = torch.empty(N, R, device="cuda")
                check(lib.op3_decoder_bwd(C.byref(d), N, ptr(prep), ptr(z), ptr(acts), ptr(seed), ptr(scratch), ptr(dz), 0, None, st))
                torch.cuda.synchronize()
                res.append((out, dz))
            assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]), (K, B)
    finally:
        lib.op3_set_tcq_pair(0)
        op3_b200.set_precision("fp32")
