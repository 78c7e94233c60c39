#!/usr/bin/env python
"""Benchmark of the OP3 hot path (BASELINE.json metric: train sequences/sec at 1/2/4/8 B200; MPC candidates/sec).

    python bench.py --gpus N --steps K --warmup W [--config c2]      # this repo (sm_100a kernels); torchrun for N > 1
    python bench.py --impl reference --steps K --warmup W [--config] # the CPU oracle port of the reference, host cores

--config selects one of BASELINE.json's configurations (default c2, the one the headline metric is quoted on):
  c1            stack variant, B=8, K=7, T=2, schedule 00001 (the reference's own CPU-runnable case)
  c2            pickplace variant, B=64/GPU, K=4, T=5, A=4, schedule 000011110
  c3            cloth variant (same architecture as c2, cloth action range), B=64/GPU -- the 2/4/8-GPU data-parallel case
  c4            scaled: 128x128 frames, K=14, B=64/GPU (512 on 8 GPUs), schedule 00001
  c5-pickplace  MPC: 4096 candidate actions out of one B=1 state through dynamics+decoder, Cost(final_recon, sum), sort;
                candidates sharded across ranks, costs all-gathered
  c5-stack      MPC: 4096 candidate after-images, schedule 00001 (K=7), Cost(subimage, min) against 7 goal latents

One "step" = one trainer step (op3_trainer.py:166-197: forward over the unrolled schedule, backward, gradient all-reduce
when N > 1, clip_grad_norm_(5.0) + Adam) on one batch of synthetic frames, or one CEM scoring pass over the 4096
candidates (mpc_pickplace.py:321-340 / mpc_stack.py:249-258).  Prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c1": dict(kind="train", B=8, K=7, T=2, A=0, D=64, schedule=[0, 0, 0, 0, 1], actions=None, cpu_B=8,
               workload="OP3 stack variant train step: B=8, K=7, T=2, 64x64, schedule 00001, loss weights 1..5, "
                        "random-init weights"),
    "c2": dict(kind="train", B=64, K=4, T=5, A=4, D=64, schedule=[0, 0, 0, 0, 1, 1, 1, 1, 0], actions="pickplace", cpu_B=4,
               workload="OP3 pickplace variant train step: B=64/GPU, K=4, T=5, A=4, 64x64, schedule 000011110, "
                        "loss weights 1..9, random-init weights"),
    "c3": dict(kind="train", B=64, K=4, T=5, A=4, D=64, schedule=[0, 0, 0, 0, 1, 1, 1, 1, 0], actions="cloth", cpu_B=4,
               workload="OP3 cloth variant train step: B=64/GPU, K=4, T=5 window of the dataset's 20-frame clips, A=4 in "
                        "[-1,1], 64x64, schedule 000011110, loss weights 1..9, random-init weights"),
    "c4": dict(kind="train", B=64, K=14, T=2, A=0, D=128, schedule=[0, 0, 0, 0, 1], actions=None, cpu_B=1,
               workload="scaled OP3 train step: 128x128 frames, K=14 (stack doubled), B=64/GPU (512 on 8 GPUs), T=2, "
                        "schedule 00001, loss weights 1..5, random-init weights"),
    "c5-pickplace": dict(kind="mpc", Na=4096, K=4, A=4, D=64, schedule=[1], cpu_Na=32,
                         workload="OP3 pickplace MPC scoring pass: 4096 candidate actions from one B=1 state (3 seed "
                                  "frames) through dynamics + decoder, Cost(final_recon, mse, sum), sort; candidates "
                                  "sharded across GPUs"),
    "c5-stack": dict(kind="mpc", Na=4096, K=7, A=0, D=64, schedule=[0, 0, 0, 0, 1], cpu_Na=4,
                     workload="OP3 stack MPC scoring pass: 4096 candidate after-images (T=1), schedule 00001, K=7, "
                              "Cost(subimage, mse, min) against 7 goal latents, sort; candidates sharded across GPUs"),
}


def config_dict(name):
    """The `config` object of the JSON line: identical in both arms (ours / --impl reference)."""
    c = CONFIGS[name]
    d = {"workload": c["workload"], "name": name, "K": c["K"], "image": c["D"], "schedule": "".join(map(str, c["schedule"]))}
    if c["kind"] == "train":
        d.update(batch_per_gpu=c["B"], T=c["T"],
                 l2="per-step working set (GBs of activations) >> 126 MB L2; no explicit flush")
    else:
        d.update(candidates=c["Na"], l2="per-pass working set (>= 1 GB of decoder activations) >> 126 MB L2; no explicit flush")
    return d


def synth_batch(cfg, B, seed):
    g = torch.Generator().manual_seed(seed)
    D = cfg["D"]
    frames = torch.randint(0, 256, (B, cfg["T"], 3, D, D), generator=g, dtype=torch.uint8)
    if cfg["actions"] == "pickplace":          # env bounds block_pick_and_place.py:58 (x, y of pick and place)
        lo, hi = torch.tensor([-2.5, 1.0, -2.5, 1.0]), torch.tensor([2.5, 4.0, 2.5, 4.0])
        actions = lo + (hi - lo) * torch.rand(B, cfg["T"], 4, generator=g)
    elif cfg["actions"] == "cloth":
        actions = 2.0 * torch.rand(B, cfg["T"], 4, generator=g) - 1.0
    else:
        actions = None
    return frames, actions


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], d.get("bf16_tflops_sustained", d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout.readlines()), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle port = the reference's algorithm on PyTorch CPU ops; the only places bench.py executes oracle/)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_train_rate(cfg, B, steps, warmup, threads):
    """sequences/sec of the same train step at batch B on the host cores."""
    from oracle import op3_oracle as orc
    torch.set_num_threads(threads)
    p = orc.make_params(cfg["A"], seed=0)
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v2 = {k: torch.zeros_like(v) for k, v in p.items()}
    sched, K, D = cfg["schedule"], cfg["K"], cfg["D"]
    lw = np.arange(1, len(sched) + 1)
    times = []
    for i in range(warmup + steps):
        frames, actions = synth_batch(cfg, B, 100 + i)
        t0 = time.time()
        q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        out = orc.run_schedule(q, frames.float() / 255.0, actions, sched, lw,
                               orc.make_noise(B, K, 64, len(sched) + 1, seed=1000 + 20 * i), K, D=D)
        out["total_loss"].backward()
        grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in q.items()}
        orc.clip_adam_step(p, grads, m, v2, step=i + 1)
        if i >= warmup:
            times.append(time.time() - t0)
    return B * len(times) / sum(times), float(np.mean(times)) * 1e3


def cpu_mpc_rate(cfg, Na, steps, warmup, threads):
    """candidates/sec of the same scoring pass over Na candidates on the host cores."""
    from oracle import op3_oracle as orc
    torch.set_num_threads(threads)
    K, A, sched = cfg["K"], cfg["A"], cfg["schedule"]
    p = orc.make_params(A, seed=0)
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}       # refine() differentiates the decoder
    g = torch.Generator().manual_seed(200)
    state = None
    if A:        # pickplace: the B=1 state from 3 seed frames (outside the timed region, as in bench's GPU leg)
        s0 = orc.run_schedule(q, torch.rand(1, 3, 3, 64, 64, generator=g), torch.rand(1, 3, 4, generator=g),
                              [0, 0, 1, 0, 1, 0], [0] * 6, orc.make_noise(1, K, 64, 7, seed=300), K)["state"]
        rep = lambda t: None if t is None else t.detach().repeat(Na, *([1] * (t.dim() - 1)))
        state = {"prior": {"lambdas1": rep(s0["prior"]["lambdas1"]), "lambdas2": rep(s0["prior"]["lambdas1"])},
                 "post": {k: rep(v) for k, v in s0["post"].items() if k != "extra_info"}}
        state["post"]["extra_info"] = [s0["post"]["extra_info"][0].detach().repeat(Na, 1),
                                       s0["post"]["extra_info"][1].detach().repeat(Na, 1)]
    times = []
    for i in range(warmup + steps):
        t0 = time.time()
        if A:        # one dynamics step + decode per candidate action, cost against the goal image
            cand = torch.rand(Na, 1, 4, generator=g)
            with torch.no_grad():
                o = orc.run_schedule(p, None, cand, sched, [0], orc.make_noise(Na, K, 64, 1, seed=400 + i), K, init_state=state)
                goal = torch.rand(1, 3, 64, 64, generator=g)
                orc.mpc_cost(goal, goal, o["colors_list"][:, -1] * o["masks_list"][:, -1], o["final_recon"], "final_recon", "sum")
        else:        # stack: candidate after-images, 4 refinements + 1 dynamics step, cost against K goal latents
            obs = torch.rand(Na, 1, 3, 64, 64, generator=g)
            o = orc.run_schedule(q, obs, None, sched, [0] * len(sched), orc.make_noise(Na, K, 64, len(sched) + 1, seed=400 + i), K)
            goal = torch.rand(K, 3, 64, 64, generator=g)
            with torch.no_grad():
                orc.mpc_cost(goal, goal[:1], (o["colors_list"][:, -1] * o["masks_list"][:, -1]).detach(), o["final_recon"].detach(),
                             "subimage", "min")
        if i >= warmup:
            times.append(time.time() - t0)
    return Na * len(times) / sum(times), float(np.mean(times)) * 1e3


def metric_of(cfg):
    if cfg["kind"] == "train":
        return "train_sequences_per_sec", "sequences/s", "weak"
    return "mpc_candidates_per_sec", "candidates/s", "strong"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    threads = os.cpu_count() or 1
    metric, unit, scaling = metric_of(cfg)
    if cfg["kind"] == "train":
        Bs = cfg["cpu_B"]
        rate, ms = cpu_train_rate(cfg, Bs, args.steps, args.warmup, threads)
        sample = ("each step = the same train step (same schedule, K, T, image size, clip+Adam) on a bounded sample of B=%d "
                  "sequences instead of %d per GPU (oracle/op3_oracle.py, PyTorch CPU ops, closed-form refinement inputs)"
                  % (Bs, cfg["B"]))
    else:
        Bs = cfg["cpu_Na"]
        rate, ms = cpu_mpc_rate(cfg, Bs, args.steps, args.warmup, threads)
        sample = ("each step = the same scoring pass on a bounded sample of %d candidates instead of %d "
                  "(oracle/op3_oracle.py, PyTorch CPU ops)" % (Bs, cfg["Na"]))
    line = {"impl": "reference", "metric": metric, "value": rate, "unit": unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args.config),
            "sample_units_per_step": Bs,
            "cpu_baseline": {"value": rate, "unit": unit, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# rooflines of the mixture kernels (HBM)
# ----------------------------------------------------------------------------------------------------------------------
def ncu_traffic(kernel, precision):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the family's main kernel, from the committed
    `ncu --set full` captures (profiles/r02_ncu_traffic.json, else round 1's); None when that kernel/mode was not captured."""
    for name in ("r02_ncu_traffic.json", "r01_ncu_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                v = json.load(f).get("%s/%s" % (kernel, precision), {}).get("dram_bytes_per_launch")
            if v is not None:
                return v
        except (OSError, ValueError):
            pass
    return None


def _mixture_sets(model, dev, B, K, D, n_sets, with_bwd):
    from op3_b200 import _lib
    from op3_b200._lib import Op3MixtureIO, check, ptr
    lib = _lib.load()
    eng = model.engine
    N = B * K
    d = eng.dims(B)
    ps = eng.params_struct()
    st = torch.cuda.current_stream(dev).cuda_stream
    g = torch.Generator(device=dev).manual_seed(3)
    f = lambda *s: torch.rand(*s, device=dev, generator=g)
    sets = []
    for _ in range(n_sets):
        keep = dict(dec=f(N, D, D, 4), img=f(B, 3, D, D), l=[f(B, K, 64) for _ in range(4)],
                    stats=torch.empty(lib.op3_mixture_stats_floats(C.byref(d)), device=dev), losses=torch.zeros(4, device=dev),
                    mix1=torch.empty(N, D, D, 4, device=dev), mix2=torch.empty(N, D, D, 4, device=dev),
                    smp=torch.empty(B, D, D, 4, device=dev), seed=torch.empty(N, D, D, 4, device=dev))
        io = Op3MixtureIO()
        io.dec_out, io.image, io.image_bstride = ptr(keep["dec"]), ptr(keep["img"]), 3 * D * D
        io.post_l1, io.post_l2, io.prior_l1, io.prior_l2 = [ptr(t) for t in keep["l"]]
        io.stats, io.losses = ptr(keep["stats"]), ptr(keep["losses"])
        io.mix1, io.mix2, io.smp, io.seed = ptr(keep["mix1"]), ptr(keep["mix2"]), ptr(keep["smp"]), ptr(keep["seed"])
        io.sync = eng.mix_sync(dev, st).data_ptr()
        if with_bwd:
            keep["d_in"] = f(N, 5, D, D, 4) - 0.5
            keep["d_dec"] = torch.empty(N, D, D, 4, device=dev)
            check(lib.op3_mixture_fwd(C.byref(d), C.byref(ps), C.byref(io), 1, st), "mixture_fwd")
        sets.append((keep, io))
    return lib, d, ps, st, sets


def mixture_roofline(model, dev, cfg, hbm_peak, which, backward):
    """Achieved HBM GB/s of the fused mixture forward / backward on the workload's shapes, timed alone with CUDA events;
    calls rotate over buffer sets that together exceed the L2 several times, so inputs always come from HBM."""
    from op3_b200._lib import check, ptr
    B, K, D = cfg["B"], cfg["K"], cfg["D"]
    per_set = (11 * K + 4 + (20 * K if backward else 0)) * 4.0 * B * D * D
    n_sets = int(max(4, min(12, np.ceil(4 * 126e6 / per_set))))
    lib, d, ps, st, sets = _mixture_sets(model, dev, B, K, D, n_sets, backward)
    gflat = torch.zeros_like(model.engine.flat_params())
    gs = model.engine._struct_for(gflat)
    n_it, groups = 40, []
    for g_i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_it):
            keep, io = sets[i % n_sets]
            if backward:
                check(lib.op3_mixture_bwd(C.byref(d), C.byref(io), 0.2, ptr(keep["d_in"]), ptr(keep["d_dec"]), C.byref(gs), None, st),
                      "mixture_bwd")
            else:
                check(lib.op3_mixture_fwd(C.byref(d), C.byref(ps), C.byref(io), 1, st), "mixture_fwd")
        e1.record()
        torch.cuda.synchronize(dev)
        if g_i:
            groups.append(e0.elapsed_time(e1) / n_it)
    ms = float(np.median(groups))
    alg_bytes = ((14 * K + 3) if backward else (11 * K + 4)) * 4.0 * B * D * D
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm",
            "kernel": ("mixture backward (mix_bwd_kernel + LayerNorm2D parameter-gradient reduction)" if backward
                       else "mixture forward (softmax over K, mixture likelihood, KL, LayerNorm2D, refinement inputs)"),
            "l2": "%d rotating buffer sets >> 126 MB L2" % n_sets, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak, "traffic": ncu_traffic("mixture_bwd" if backward else "mixture_fwd", "any"),
            "ms_per_call": ms, "algorithmic_bytes_per_call": alg_bytes, "peak_source": which}


# ----------------------------------------------------------------------------------------------------------------------
def build_model(cfg, dev):
    from op3_b200.torch.op3_modules import decoder_network, refinement_network
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    dec, ref = "reg", "size_dependent_conv"
    if cfg["D"] != 64:                   # new registry keys, as the reference would add them
        dec, ref = decoder_network.register_image_size(cfg["D"]), refinement_network.register_image_size(cfg["D"])
    variant = dict(refinement_model_type=ref, decoder_model_type=dec, dynamics_model_type="reg_ac32",
                   sto_repsize=64, det_repsize=64, K=cfg["K"], extra_args=dict(beta=1e-2, deterministic_sampling=False))
    # random-init weights (the reference's own initialisers); the same seed on every rank: identical replicas
    torch.manual_seed(0)
    model = create_model_v2(variant, 64, 64, cfg["A"])
    return model.to(dev)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default=os.environ.get("OP3_BENCH_CONFIG", "c2"), choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch / candidate count of the config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of one CUDA graph")
    ap.add_argument("--no-wgrad-overlap", action="store_true",
                    help="conv weight gradients on the launch stream instead of the library's side stream (A/B)")
    ap.add_argument("--precision", default=os.environ.get("OP3_PRECISION", "f16x3"), choices=["fp32", "tf32x3", "tf32", "f16x3"],
                    help="arithmetic of the 5x5 convolutions (see op3_set_precision)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3                                            # timing rule: at least 3 warm-up steps
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["B" if cfg["kind"] == "train" else "Na"] = args.batch
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the OP3 kernels are sm_100a only (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation: keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    from op3_b200 import _lib
    lib = _lib.load()
    import op3_b200
    op3_b200.set_precision(args.precision)
    model = build_model(cfg, dev)
    torch.manual_seed(1234 + rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    debug_steps = bool(os.environ.get("OP3_BENCH_DEBUG"))

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            if debug_steps:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            fn(i)
            if debug_steps:
                t1 = time.perf_counter()
                torch.cuda.synchronize()
                print("[bench debug] %s step %d: host %.1f ms, host+device %.1f ms, allocated %.2f GB, reserved %.2f GB"
                      % (fn.__name__, i, 1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t0), torch.cuda.memory_allocated() / 2**30,
                         torch.cuda.memory_reserved() / 2**30), file=sys.stderr)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    extra = {}
    if cfg["kind"] == "train":
        from op3_b200.torch.op3_modules.op3_trainer import OP3Trainer, TrainingScheduler
        B, K, T, D = cfg["B"], cfg["K"], cfg["T"], cfg["D"]
        trainer = OP3Trainer(None, None, model, TrainingScheduler(4, "curriculum", T, T), B, lr=3e-4)
        # the captured step keeps its activations in a private pool next to the eager pool of the profiling pass: at the
        # scaled config (~85 GB per step) only one of them fits in HBM, and launch gaps are invisible at 1 s per step
        use_graph = (not args.no_graph) and cfg["D"] == 64
        trainer.use_cuda_graph = use_graph
        model.engine.wgrad_overlap = not args.no_wgrad_overlap
        extra["wgrad_overlap"] = bool(model.engine.wgrad_overlap)
        sched = np.array(cfg["schedule"], dtype=np.float64)
        lw = np.arange(1, len(sched) + 1)
        n_batches = 4
        host = []
        for i in range(n_batches):
            fr, ac = synth_batch(cfg, B, 17 + 31 * rank + i)
            host.append((fr.pin_memory(), None if ac is None else ac.pin_memory()))
        resident = [(fr.to(dev).float() / 255.0, None if ac is None else ac.to(dev)) for fr, ac in host]
        out4 = torch.empty(4, pin_memory=True)

        def step_resident(i):
            fr, ac = resident[i % n_batches]
            return trainer.train_step(fr, ac, sched, lw)

        from op3_b200.torch.op3_modules.op3_trainer import ScalarReadback
        readback = ScalarReadback(dev, 4, trainer.stat_lag)
        e2e_rows = []

        def step_e2e(i):
            # the loop body of OP3Trainer.train_epoch: every step copies its batch from pinned host memory, and every step's
            # four scalars travel back to the host -- read one step later (stat_lag = 1), while the next step runs
            fr, ac = host[i % n_batches]
            frd, acd = trainer.prepare_tensors((fr, ac))      # pinned uint8 -> H2D as bytes -> op3_prepare_frames (/255)
            vals = trainer.train_step(frd, acd, sched, lw)
            e2e_rows.extend(readback.push(vals))
            return vals

        def step_e2e_sync(i):
            # the reference's own loop: loss.item() after every batch (op3_trainer.py:203-213) -- the GPU waits for the host
            fr, ac = host[i % n_batches]
            frd, acd = trainer.prepare_tensors((fr, ac))
            vals = trainer.train_step(frd, acd, sched, lw)
            out4.copy_(torch.stack(vals), non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            return vals
        units_per_step = B * world
        h2d = int(B * T * 3 * D * D + (B * T * 4 * 4 if cfg["A"] else 0))
        d2h = 16
        extra["cuda_graph"] = bool(trainer.use_cuda_graph)
    else:
        from op3_b200.mpc import Cost
        from op3_b200 import parallel as par
        from op3_b200.torch.op3_modules.op3_trainer import frames_to_float
        Na, K, D = cfg["Na"], cfg["K"], cfg["D"]
        model.eval_mode = True
        model.inference_chunk = 512
        g = torch.Generator().manual_seed(31)                  # same candidates / state on every rank
        sched = np.array(cfg["schedule"], dtype=np.float64)
        if cfg["A"]:       # pickplace (mpc_pickplace.py:253-340): state from 3 seed frames, candidates = actions
            seed_fr = (torch.randint(0, 256, (1, 3, 3, D, D), generator=g, dtype=torch.uint8).float() / 255.0).to(dev)
            seed_ac = (2 * torch.rand(1, 3, 4, generator=g) - 1).to(dev)
            info = model.batch_internal_inference(seed_fr, seed_ac, None, np.array([0., 0., 1., 0., 1., 0.]))
            state0 = info["state"]
            lo, hi = torch.tensor([-2.5, 1.0, -2.5, 1.0]), torch.tensor([2.5, 4.0, 2.5, 4.0])
            host = [(lo + (hi - lo) * torch.rand(Na, 1, 4, generator=g)).pin_memory() for _ in range(4)]
            resident = [h.to(dev) for h in host]
            goal_info = {"goal_image": (torch.randint(0, 256, (1, 3, D, D), generator=g, dtype=torch.uint8).float() / 255.0).to(dev)}
            cost = Cost(None, core_type="final_recon", compare_func="mse", post_process="raw", aggregate="sum")
            h2d = Na * 4 * 4

            def score(cand):
                return par.sharded_rollout_costs(model, cost, goal_info, cand, state0, sched)
            to_dev = lambda h: h.to(dev, non_blocking=True)
        else:              # stack (mpc_stack.py:249-258, 403-423): goal latents from 5 refinements, candidates = images
            goal_fr = (torch.randint(0, 256, (1, 1, 3, D, D), generator=g, dtype=torch.uint8).float() / 255.0).to(dev)
            ginfo = model.batch_internal_inference(goal_fr, None, None, np.zeros(5))
            goal_info = {"sub_images": ginfo["sub_images"], "goal_image": goal_fr[:, 0]}
            host = [torch.randint(0, 256, (Na, 1, D, D, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(2)]
            resident = [frames_to_float(h.to(dev)) for h in host]
            cost = Cost(None, core_type="subimage", compare_func="mse", post_process="raw", aggregate="min")
            h2d = Na * D * D * 3

            def score(cand):
                return par.sharded_rollout_costs(model, cost, goal_info, None, None, sched, obs=cand)
            to_dev = lambda h: frames_to_float(h.to(dev, non_blocking=True))
        out2 = torch.empty(2, pin_memory=True)

        def step_resident(i):
            return score(resident[i % len(resident)])

        def step_e2e(i):
            sorted_costs, order = score(to_dev(host[i % len(host)]))
            out2.copy_(torch.stack([sorted_costs[0], order[0].float()]), non_blocking=True)   # selected action + its cost
            torch.cuda.current_stream(dev).synchronize()
        units_per_step = Na
        d2h = 8
        if world > 1:      # the exchange step on its own: all-gather of the Na fp32 costs
            s, e = par.shard_range(Na)
            loc = torch.rand(e - s, device=dev)
            for _ in range(3):
                par.all_gather_rows(loc, Na)
            extra["allgather_ms"] = timed(lambda i: par.all_gather_rows(loc, Na), 20) / 20

    for i in range(args.warmup + (2 if cfg["kind"] == "mpc" else 0)):   # the MPC passes also warm the caching allocator
        step_resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = lib.op3_launch_count()
    ms_total = timed(step_resident, args.steps)                  # the headline: no instrumentation inside the timed region
    launches = lib.op3_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # second pass over the same K steps with the per-family CUDA events switched on (they cost ~2 ms per step and cannot
    # be recorded inside a CUDA graph, which is why the headline is timed without them): kernel-family times and the
    # roofline numerators come from here, with every kernel launched from the host
    if cfg["kind"] == "train":
        trainer.use_cuda_graph = False
        model.engine.wgrad_overlap = False     # per-family times are kernel durations: no two families share the SMs here
    lib.op3_profile_enable(1)
    n_prof = min(args.steps, 5)
    l1 = lib.op3_launch_count()
    ms_prof = timed(step_resident, n_prof)
    launches_host = (lib.op3_launch_count() - l1) / n_prof
    pms = (C.c_double * 4)(); pwork = (C.c_double * 4)(); pn = (C.c_longlong * 4)()
    lib.op3_profile_collect(pms, pwork, pn)
    lib.op3_profile_enable(0)
    if cfg["kind"] == "train":
        trainer.use_cuda_graph = use_graph
        model.engine.wgrad_overlap = not args.no_wgrad_overlap
    # end to end: pinned host inputs in, result scalars out, every step
    for i in range(min(2, args.warmup)):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    ms_e2e_sync = None
    if cfg["kind"] == "train":
        e2e_rows.extend(readback.drain())
        assert len(e2e_rows) == args.steps + min(2, args.warmup) and all(np.isfinite(r).all() for r in e2e_rows), "e2e read-back"
        ms_e2e_sync = timed(step_e2e_sync, args.steps)

    units = units_per_step * args.steps
    value = units / (ms_total * 1e-3)
    e2e_value = units / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    metric, unit, scaling = metric_of(cfg)
    graphed = cfg["kind"] == "train" and use_graph
    hbm, bf16, bf16_sus, which = peaks()
    tf32_peak = bf16_sus / 2.0           # dense TF32 tensor-pipe peak ~ 1/2 of the measured (sustained) bf16 figure
    # the tensor-pipe peak of the mode's operand type (fp16 runs at the bf16 rate), and how many tensor products the mode
    # issues per algorithmic FLOP (hi*hi + hi*lo + lo*hi)
    mode_peak = {"fp32": tf32_peak, "tf32x3": tf32_peak, "tf32": tf32_peak, "f16x3": bf16_sus}[args.precision]
    mode_products = {"fp32": 1.0, "tf32x3": 3.0, "tf32": 1.0, "f16x3": 3.0}[args.precision]
    kind = {"fp32": "fp32 CUDA cores", "tf32x3": "tcgen05 kind::tf32, 3xTF32 split", "tf32": "tcgen05 kind::tf32",
            "f16x3": "tcgen05 kind::f16, fp16 hi/lo split"}[args.precision]
    fam = ["conv5x5 forward/dgrad", "conv5x5 wgrad", "mixture", "small GEMM"]
    fams = {fam[i]: {"ms_per_step": pms[i] / n_prof, "calls_per_step": pn[i] / n_prof,
                     "tflops": (pwork[i] / (pms[i] * 1e-3) / 1e12) if (pms[i] > 0 and i != 2) else None}
            for i in range(4)}
    dom = 0 if pms[0] >= pms[1] else 1
    achieved = pwork[dom] / (pms[dom] * 1e-3) / 1e12 if pms[dom] > 0 else 0.0
    line = {
        "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "f32 (3xTF32 split on tcgen05)", "tf32": "tf32",
                                       "f16x3": "f32 (fp16 hi/lo operand split on tcgen05, fp32 accumulate)"}[args.precision],
        "data": "synthetic", "config": config_dict(args.config), "precision": args.precision,
        "e2e": dict({"value": e2e_value, "unit": unit, "ms_per_step": ms_e2e / args.steps,
                     "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                    **({"readback": "every step's scalars copied to pinned host memory and read one step later "
                                    "(OP3Trainer.train_epoch, stat_lag=1)",
                        "sync_value": units / (ms_e2e_sync * 1e-3), "sync_ms_per_step": ms_e2e_sync / args.steps,
                        "sync_readback": "blocking read after every step (the reference's loss.item())"}
                       if ms_e2e_sync is not None else {})),
        # kernels of this library that EXECUTE inside the timed region.  Under CUDA-graph replay the host issues only the
        # graph launch + the optimizer's kernels (host_launches); the graph's kernel nodes are the same launches the
        # uncaptured step issues, counted by the library's launch counter in the profiling pass
        "gpu_launches": int(round(launches_host * args.steps)) if graphed else int(launches),
        "gpu_launches_per_step": launches_host if graphed else launches / args.steps,
        "host_launches": int(launches), "cuda_graph": bool(graphed), "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "%s (%s)" % (fam[dom], kind),
                     "achieved": achieved, "peak": mode_peak, "unit": "TFLOP/s", "frac": achieved / mode_peak,
                     "traffic": ncu_traffic(("conv_fwd", "conv_wgrad")[dom], args.precision),
                     "avg_launch_ms": pms[dom] / max(pn[dom], 1),
                     "share_of_step": pms[dom] / ms_prof,
                     "peak_source": ("%s bf16_tflops_sustained (dense fp16/bf16)" if args.precision == "f16x3"
                                     else "%s bf16_tflops_sustained / 2 (dense TF32)") % which,
                     # the split modes issue three tensor-core products per algorithmic FLOP: their own ceiling is peak / 3
                     "frac_of_mode_ceiling": achieved / (mode_peak / mode_products)},
        "kernel_families": fams,
    }
    if args.batch:
        line["config"]["batch_override"] = args.batch
    line.update(extra)
    if cfg["kind"] == "train":
        line["roofline_mixture"] = mixture_roofline(model, dev, cfg, hbm, which, backward=False)
        line["roofline_mixture_bwd"] = mixture_roofline(model, dev, cfg, hbm, which, backward=True)
    if not args.no_cpu_baseline and world == 1:       # the CPU oracle port is timed at N = 1 only (rank 0, host cores)
        threads = os.cpu_count() or 1
        if cfg["kind"] == "train":
            Bs = cfg["cpu_B"] if args.config == "c1" else min(cfg["cpu_B"], 4)
            rate, ms = cpu_train_rate(cfg, Bs, 2 if cfg["D"] == 64 else 1, 0, threads)
            sample = "%d step(s) of B=%d sequences of the same workload (oracle port, PyTorch CPU)" % (2 if cfg["D"] == 64 else 1, Bs)
        else:
            rate, ms = cpu_mpc_rate(cfg, cfg["cpu_Na"], 1, 0, threads)
            sample = "1 scoring pass over %d candidates of the same workload (oracle port, PyTorch CPU)" % cfg["cpu_Na"]
        line["cpu_baseline"] = {"value": rate, "unit": unit, "cores": threads, "kind": "port", "sample": sample, "ms_per_step": ms}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
