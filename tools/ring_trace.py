"""Debug: timeline of the ring wgrad issuer (needs a build with OP3_NVCC_EXTRA=-DOP3_TC_TRACE).  Runs one train step of the
bench workload and prints where the issuer thread of conv5x5_wgrad_ring_kernel spent its cycles (all launches)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import op3_b200
    from op3_b200 import _lib
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    from op3_b200.torch.op3_modules.op3_trainer import OP3Trainer, TrainingScheduler
    lib = _lib.load()
    op3_b200.set_precision("f16x3")
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    variant = dict(refinement_model_type="size_dependent_conv", decoder_model_type="reg", dynamics_model_type="reg_ac32",
                   sto_repsize=64, det_repsize=64, K=4, extra_args=dict(beta=1e-2, deterministic_sampling=False))
    model = create_model_v2(variant, 64, 64, 4).to(dev)
    trainer = OP3Trainer(None, None, model, TrainingScheduler(4, "curriculum", 5, 5), 64, lr=3e-4)
    cfg = bench.CONFIGS["c2"]
    fr, ac = bench.synth_batch(cfg, 64, 1)
    fr, ac = fr.to(dev).float() / 255.0, ac.to(dev)
    sched, lw = np.array(cfg["schedule"], dtype=np.float64), np.arange(1, 10)
    for _ in range(2):
        trainer.train_step(fr, ac, sched, lw)
    torch.cuda.synchronize()
    for name in ("op3_debug_ring_trace", "op3_debug_tc_trace"):
        f = getattr(lib, name)
        f.argtypes = [C.c_void_p, C.c_int]
        f(None, 1)
    trainer.train_step(fr, ac, sched, lw)
    torch.cuda.synchronize()
    buf = np.zeros((256, 8), dtype=np.uint64)
    lib.op3_debug_ring_trace(buf.ctypes.data, 0)
    full = buf[:148].astype(np.float64)
    t = full.mean(0)
    hs = t[3] / 2
    print("ring G loader (warp 2 of group 0, per step of its own): convert %.0f, wait for empty %.0f, stores %.0f, proxy fence %.0f cycles"
          % (t[5] / hs, t[6] / hs, t[7] / hs, t[4] / hs))
    print("ring wgrad, one step, mean over CTAs: wait operands %.0f cyc of %.0f total (%.1f%%), MMAs %.0f, steps %.0f"
          % (t[0], t[1], 100 * t[0] / t[1], t[2], t[3]))
    print("  cycles per MMA excluding operand waits: %.1f; including: %.1f" % ((t[1] - t[0]) / t[2], t[1] / t[2]))
    buf = np.zeros((256, 4), dtype=np.uint64)
    lib.op3_debug_tc_trace(buf.ctypes.data, 0)
    t = buf[:148].astype(np.float64).mean(0)
    print("conv fwd/dgrad, one step: wait operands %.0f, wait accumulators %.0f, total %.0f cyc, MMAs %.0f -> %.1f cyc per MMA pair issuing"
          % (t[0], t[1], t[2], t[3], (t[2] - t[0] - t[1]) / (t[3] / 2)))


if __name__ == "__main__":
    main()
