"""Times (or, under ncu, exposes) the mixture forward / backward alone on a bench config's shapes (default c2).
usage: python tools/mix_probe.py [config]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    cfg = dict(bench.CONFIGS[name])
    from op3_b200.torch.op3_modules import decoder_network, refinement_network
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    variant = dict(refinement_model_type="size_dependent_conv", decoder_model_type="reg", dynamics_model_type="reg_ac32",
                   sto_repsize=64, det_repsize=64, K=cfg["K"], extra_args=dict(beta=1e-2, deterministic_sampling=False))
    if cfg["D"] != 64:
        variant["decoder_model_type"] = decoder_network.register_image_size(cfg["D"])
        variant["refinement_model_type"] = refinement_network.register_image_size(cfg["D"])
    model = create_model_v2(variant, 64, 64, cfg["A"]).to(dev)
    hbm = bench.peaks()[0]
    import ctypes as C
    from op3_b200 import _lib
    lib = _lib.load()
    trace = getattr(lib, "op3_debug_mix_trace", None) if hasattr(C.CDLL(_lib.LIB_PATH), "op3_debug_mix_trace") else None
    if trace is not None:
        trace.argtypes = [C.c_void_p, C.c_int]
        trace(None, 1)
    for bwd in (False, True):
        r = bench.mixture_roofline(model, dev, cfg, hbm, "measured", backward=bwd)
        print("%s: %.1f us per call, %.0f GB/s, %.3f of HBM" % ("backward" if bwd else "forward ", 1e3 * r["ms_per_call"], r["achieved"], r["frac"]))
        if trace is not None and not bwd:
            buf = (C.c_ulonglong * 16)()
            trace(buf, 1)
            n = max(1, buf[15])
            names = ["prologue", "wait first sub-chunk", "pass 1", "CTA reduction", "cluster barrier", "statistics", "syncthreads", "pass 2", "exit barrier"]
            print("  cycles per CTA (thread 0): " + " | ".join("%s %d" % (nm, buf[i] // n) for i, nm in enumerate(names)))


if __name__ == "__main__":
    main()
