"""Times (or, under ncu, exposes) the mixture forward / backward alone on the bench workload's shapes.
usage: python tools/mix_probe.py [B K D]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    a = [int(x) for x in sys.argv[1:4]]
    if len(a) == 3:
        bench.B_PER_GPU, bench.K, bench.D = a
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    import op3_b200
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    variant = dict(refinement_model_type="size_dependent_conv", decoder_model_type="reg", dynamics_model_type="reg_ac32",
                   sto_repsize=64, det_repsize=64, K=bench.K, extra_args=dict(beta=1e-2, deterministic_sampling=False))
    model = create_model_v2(variant, 64, 64, 4)
    model.image_size = bench.D
    model.to(dev)
    hbm = bench.peaks()[0]
    print(bench.mixture_roofline(model, dev, hbm, "measured"))
    print(bench.mixture_bwd_roofline(model, dev, hbm, "measured"))


if __name__ == "__main__":
    main()
