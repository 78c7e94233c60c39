"""One trainer step of the bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import op3_b200
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    from op3_b200.torch.op3_modules.op3_trainer import OP3Trainer, TrainingScheduler
    op3_b200.set_precision(sys.argv[1] if len(sys.argv) > 1 else "f16x3")
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
    torch.cuda.cudart().cudaProfilerStart()
    trainer.train_step(fr, ac, sched, lw)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
