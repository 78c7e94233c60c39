#!/usr/bin/env python
"""Benchmark of the OP3 training step (BASELINE.json metric: train sequences/sec at 1/2/4/8 B200).

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels); torchrun for N > 1
    python bench.py --impl reference --steps K --warmup W    # the CPU oracle port of the reference, host cores

One "step" = one trainer step (op3_trainer.py:166-197: forward over the unrolled schedule, backward, gradient
all-reduce when N > 1, clip_grad_norm_(5.0) + Adam) on one batch of synthetic frames of the 'pickplace' variant
(BASELINE.json configs[1]: B=64 per GPU, K=4, T=5 frames, A=4, schedule [0,0,0,0,1,1,1,1,0]).  Prints ONE JSON line.
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

SCHEDULE = [0, 0, 0, 0, 1, 1, 1, 1, 0]
B_PER_GPU, K, T, A, D = 64, 4, 5, 4, 64
CONFIG = {"workload": "OP3 pickplace variant train step: B=64/GPU, K=4, T=5, A=4, 64x64, schedule 000011110, "
                      "loss weights 1..9, random-init weights", "batch_per_gpu": B_PER_GPU, "K": K, "T": T,
          "image": D, "schedule": "".join(map(str, SCHEDULE)),
          "l2": "per-step working set (~8 GB of activations) >> 126 MB L2; no explicit flush"}


def synth_batch(B, seed):
    g = torch.Generator().manual_seed(seed)
    frames = torch.randint(0, 256, (B, T, 3, D, D), generator=g, dtype=torch.uint8)
    lo = torch.tensor([-2.5, 1.0, -2.5, 1.0])
    hi = torch.tensor([2.5, 4.0, 2.5, 4.0])
    actions = lo + (hi - lo) * torch.rand(B, T, 4, generator=g)
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


def cpu_oracle_rate(B, steps, warmup, threads):
    """Reference algorithm (oracle port) on the host cores: sequences/sec of the same train step at batch B."""
    from oracle import op3_oracle as orc
    torch.set_num_threads(threads)
    p = orc.make_params(A, seed=0)
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v2 = {k: torch.zeros_like(v) for k, v in p.items()}
    lw = np.arange(1, len(SCHEDULE) + 1)
    times = []
    for i in range(warmup + steps):
        frames, actions = synth_batch(B, 100 + i)
        t0 = time.time()
        out, grads = orc.train_step(p, frames.float() / 255.0, actions, SCHEDULE, lw,
                                    orc.make_noise(B, K, 64, len(SCHEDULE) + 1, seed=1000 + 20 * i), K)
        orc.clip_adam_step(p, grads, m, v2, step=i + 1)
        if i >= warmup:
            times.append(time.time() - t0)
    return B * len(times) / sum(times), float(np.mean(times)) * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    Bs = 2
    rate, ms = cpu_oracle_rate(Bs, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": "train_sequences_per_sec", "value": rate, "unit": "sequences/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "cpu_baseline": {"value": rate, "unit": "sequences/s", "cores": threads, "kind": "port",
                             "sample": "B=%d sequences per step of the same workload (oracle/op3_oracle.py, PyTorch CPU)" % Bs},
            "e2e": {"value": rate, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def ncu_traffic(kernel, precision):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the family's main kernel, from the committed
    `ncu --set full` capture (profiles/r01_ncu_traffic.json); None when that kernel/mode was not captured."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_traffic.json")) as f:
            return json.load(f).get("%s/%s" % (kernel, precision), {}).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def mixture_roofline(model, dev, hbm_peak, which):
    """Achieved HBM GB/s of the fused mixture forward (statistics + emit kernels) on the workload's shapes, timed alone
    with CUDA events per call; calls rotate over 4 buffer sets (300 MB >> L2), so inputs always come from HBM."""
    from op3_b200 import _lib
    from op3_b200._lib import Op3MixtureIO, check, ptr
    lib = _lib.load()
    eng = model.engine
    B, N = B_PER_GPU, B_PER_GPU * K
    d = eng.dims(B)
    ps = eng.params_struct()
    st = torch.cuda.current_stream(dev).cuda_stream
    g = torch.Generator(device=dev).manual_seed(3)
    f = lambda *s: torch.rand(*s, device=dev, generator=g)
    # R rotating buffer sets (~75 MB each: 300 MB in total >> 126 MB of L2), so every call finds its inputs in HBM
    R = 4
    sets = []
    for _ in range(R):
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
        sets.append((keep, io))
    # average launch duration: CUDA events around n_it back-to-back calls on the launch stream (median of 5 groups)
    n_it, groups = 40, []
    for g_i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_it):
            check(lib.op3_mixture_fwd(C.byref(d), C.byref(ps), C.byref(sets[i % R][1]), 1, st), "mixture_fwd")
        e1.record()
        torch.cuda.synchronize(dev)
        if g_i:
            groups.append(e0.elapsed_time(e1) / n_it)
    ms = float(np.median(groups))
    alg_bytes = (11 * K + 4) * 4.0 * B * D * D
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "mixture forward (mix_fused_kernel: one launch, 8-CTA cluster per sample)", "l2": "4 rotating buffer sets, 300 MB >> 126 MB L2", "achieved": achieved,
            "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": ncu_traffic("mixture_fwd", "any"), "ms_per_call": ms,
            "algorithmic_bytes_per_call": alg_bytes, "peak_source": which}


def mixture_bwd_roofline(model, dev, hbm_peak, which):
    """Achieved HBM GB/s of the mixture backward (d total / d dec_out from the refinement-input gradients + the loss
    weight; SURVEY.md section 8d: (14K+3)*4 bytes per sample-pixel), same rotating buffer sets as the forward."""
    from op3_b200 import _lib
    from op3_b200._lib import Op3MixtureIO, check, ptr
    lib = _lib.load()
    eng = model.engine
    B, N = B_PER_GPU, B_PER_GPU * K
    d = eng.dims(B)
    ps = eng.params_struct()
    gflat = torch.zeros_like(eng.flat_params())
    gs = eng._struct_for(gflat)
    st = torch.cuda.current_stream(dev).cuda_stream
    g = torch.Generator(device=dev).manual_seed(4)
    f = lambda *s: torch.rand(*s, device=dev, generator=g)
    R = 4
    sets = []
    for _ in range(R):
        keep = dict(dec=f(N, D, D, 4), img=f(B, 3, D, D), l=[f(B, K, 64) for _ in range(4)],
                    stats=torch.empty(lib.op3_mixture_stats_floats(C.byref(d)), device=dev), losses=torch.zeros(4, device=dev),
                    mix1=torch.empty(N, D, D, 4, device=dev), mix2=torch.empty(N, D, D, 4, device=dev),
                    smp=torch.empty(B, D, D, 4, device=dev), d_in=f(N, 5, D, D, 4) - 0.5, d_dec=torch.empty(N, D, D, 4, device=dev))
        io = Op3MixtureIO()
        io.dec_out, io.image, io.image_bstride = ptr(keep["dec"]), ptr(keep["img"]), 3 * D * D
        io.post_l1, io.post_l2, io.prior_l1, io.prior_l2 = [ptr(t) for t in keep["l"]]
        io.stats, io.losses = ptr(keep["stats"]), ptr(keep["losses"])
        io.mix1, io.mix2, io.smp = ptr(keep["mix1"]), ptr(keep["mix2"]), ptr(keep["smp"])
        io.sync = eng.mix_sync(dev, st).data_ptr()
        check(lib.op3_mixture_fwd(C.byref(d), C.byref(ps), C.byref(io), 1, st), "mixture_fwd")
        sets.append((keep, io))
    n_it, groups = 40, []
    for g_i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_it):
            keep, io = sets[i % R]
            check(lib.op3_mixture_bwd(C.byref(d), C.byref(io), 0.2, ptr(keep["d_in"]), ptr(keep["d_dec"]), C.byref(gs), None, st),
                  "mixture_bwd")
        e1.record()
        torch.cuda.synchronize(dev)
        if g_i:
            groups.append(e0.elapsed_time(e1) / n_it)
    ms = float(np.median(groups))
    alg_bytes = (14 * K + 3) * 4.0 * B * D * D
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": "mixture backward (mix_bwd_kernel + LayerNorm2D parameter-gradient reduction)",
            "l2": "4 rotating buffer sets >> 126 MB L2", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak, "traffic": ncu_traffic("mixture_bwd", "any"), "ms_per_call": ms,
            "algorithmic_bytes_per_call": alg_bytes, "peak_source": which}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default=os.environ.get("OP3_PRECISION", "f16x3"), choices=["fp32", "tf32x3", "tf32", "f16x3"],
                    help="arithmetic of the 5x5 convolutions (see op3_set_precision)")
    args = ap.parse_args()
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
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    from op3_b200.torch.op3_modules.op3_trainer import OP3Trainer, TrainingScheduler
    lib = _lib.load()
    import op3_b200
    op3_b200.set_precision(args.precision)
    # random-init weights of the pickplace architecture (the reference's own initialisers, variants.py / train_op3.py);
    # the same seed on every rank so that the data-parallel replicas start identical
    torch.manual_seed(0)
    variant = dict(refinement_model_type="size_dependent_conv", decoder_model_type="reg", dynamics_model_type="reg_ac32",
                   sto_repsize=64, det_repsize=64, K=K, extra_args=dict(beta=1e-2, deterministic_sampling=False))
    model = create_model_v2(variant, 64, 64, A)
    model.to(dev)
    torch.manual_seed(1234 + rank)
    trainer = OP3Trainer(None, None, model, TrainingScheduler(4, "curriculum", T, T), B_PER_GPU, lr=3e-4)
    sched = np.array(SCHEDULE, dtype=np.float64)
    lw = np.arange(1, len(SCHEDULE) + 1)

    n_batches = 4
    host = []
    for i in range(n_batches):
        fr, ac = synth_batch(B_PER_GPU, 17 + 31 * rank + i)
        host.append((fr.pin_memory(), ac.pin_memory()))
    resident = [(fr.to(dev).float() / 255.0, ac.to(dev)) for fr, ac in host]
    out4 = torch.empty(4, pin_memory=True)

    def step_resident(i):
        fr, ac = resident[i % n_batches]
        return trainer.train_step(fr, ac, sched, lw)

    def step_e2e(i):
        fr, ac = host[i % n_batches]
        frd, acd = trainer.prepare_tensors((fr, ac))      # pinned uint8 -> H2D as bytes -> op3_prepare_frames (/255)
        vals = trainer.train_step(frd, acd, sched, lw)
        out4.copy_(torch.stack(vals), non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return vals

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = lib.op3_launch_count()
    ms_total = timed(step_resident, args.steps)                  # the headline: no instrumentation inside the timed region
    launches = lib.op3_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    # second pass over the same K steps with the per-family CUDA events switched on (they cost ~2 ms per step, which is
    # why the headline is timed without them): kernel-family times and the roofline numerators come from here
    lib.op3_profile_enable(1)
    ms_prof = timed(step_resident, args.steps)
    pms = (C.c_double * 4)(); pwork = (C.c_double * 4)(); pn = (C.c_longlong * 4)()
    lib.op3_profile_collect(pms, pwork, pn)
    lib.op3_profile_enable(0)
    # end to end: pinned host uint8 frames + actions in, 4 loss scalars out, every step
    for i in range(min(2, args.warmup)):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)

    seqs = B_PER_GPU * world * args.steps
    value = seqs / (ms_total * 1e-3)
    e2e_value = seqs / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    hbm, bf16, bf16_sus, which = peaks()
    tf32_peak = bf16_sus / 2.0           # dense TF32 tensor-pipe peak ~ 1/2 of the measured (sustained) bf16 figure
    # the tensor-pipe peak of the mode's operand type (fp16 runs at the bf16 rate), and how many tensor products the mode
    # issues per algorithmic FLOP (hi*hi + hi*lo + lo*hi)
    mode_peak = {"fp32": tf32_peak, "tf32x3": tf32_peak, "tf32": tf32_peak, "f16x3": bf16_sus}[args.precision]
    mode_products = {"fp32": 1.0, "tf32x3": 3.0, "tf32": 1.0, "f16x3": 3.0}[args.precision]
    kind = {"fp32": "fp32 CUDA cores", "tf32x3": "tcgen05 kind::tf32, 3xTF32 split", "tf32": "tcgen05 kind::tf32",
            "f16x3": "tcgen05 kind::f16, fp16 hi/lo split"}[args.precision]
    fam = ["conv5x5 forward/dgrad", "conv5x5 wgrad", "mixture", "small GEMM"]
    fams = {fam[i]: {"ms_per_step": pms[i] / args.steps, "calls_per_step": pn[i] / args.steps,
                     "tflops": (pwork[i] / (pms[i] * 1e-3) / 1e12) if (pms[i] > 0 and i != 2) else None}
            for i in range(4)}
    dom = 0 if pms[0] >= pms[1] else 1
    achieved = pwork[dom] / (pms[dom] * 1e-3) / 1e12 if pms[dom] > 0 else 0.0
    line = {
        "metric": "train_sequences_per_sec", "value": value, "unit": "sequences/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp32": "f32", "tf32x3": "f32 (3xTF32 split on tcgen05)", "tf32": "tf32",
                                       "f16x3": "f32 (fp16 hi/lo operand split on tcgen05, fp32 accumulate)"}[args.precision],
        "data": "synthetic", "config": dict(CONFIG, precision=args.precision),
        "e2e": {"value": e2e_value, "unit": "sequences/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(B_PER_GPU * T * 3 * D * D + B_PER_GPU * T * 4 * 4), "d2h_bytes_per_step": 16},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "%s (%s)" % (fam[dom], kind),
                     "achieved": achieved, "peak": mode_peak, "unit": "TFLOP/s", "frac": achieved / mode_peak,
                     "traffic": ncu_traffic(("conv_fwd", "conv_wgrad")[dom], args.precision),
                     "avg_launch_ms": pms[dom] / max(pn[dom], 1),
                     "share_of_step": pms[dom] / ms_prof,
                     "peak_source": ("%s bf16_tflops_sustained (dense fp16/bf16)" if args.precision == "f16x3"
                                     else "%s bf16_tflops_sustained / 2 (dense TF32)") % which,
                     # the split modes issue three tensor-core products per algorithmic FLOP: their own ceiling is peak / 3
                     "frac_of_mode_ceiling": achieved / (mode_peak / mode_products)},
        "roofline_mixture": mixture_roofline(model, dev, hbm, which),
        "roofline_mixture_bwd": mixture_bwd_roofline(model, dev, hbm, which),
        "kernel_families": fams,
    }
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, ms = cpu_oracle_rate(4, 2, 0, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "sequences/s", "cores": threads, "kind": "port",
                                "sample": "2 steps of B=4 sequences of the same workload (oracle port, PyTorch CPU)",
                                "ms_per_step": ms}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
