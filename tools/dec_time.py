"""Decoder forward / latent-gradient time per call at N = B*K slot images (CUDA events, warm): A/B of kernel switches."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import op3_b200
    from op3_b200 import _lib
    from op3_b200._lib import check, ptr
    from gpu_util import build_model
    lib = _lib.load()
    op3_b200.set_precision("f16x3")
    B, K = int(os.environ.get("B", 64)), int(os.environ.get("K", 4))
    if os.environ.get("TCQ_PAIR"):
        lib.op3_set_tcq_pair(int(os.environ["TCQ_PAIR"]))
    m, _ = build_model(0, K, seed=1)
    eng = m.engine
    N, R = B * K, 128
    st = torch.cuda.current_stream().cuda_stream
    d = eng.dims(B)
    prep = eng.prepared(st)
    z = torch.randn(N, R, device="cuda")
    acts = torch.empty(lib.op3_decoder_acts_floats(C.byref(d), N), device="cuda")
    out = torch.empty(N, 64 * 64, 4, device="cuda")
    seed = torch.randn(N, 64 * 64, 4, device="cuda")
    scratch = torch.empty(lib.op3_decoder_scratch_floats(C.byref(d), N), device="cuda")
    dz = torch.empty(N, R, device="cuda")

    def fwd():
        check(lib.op3_decoder_fwd(C.byref(d), N, ptr(prep), ptr(z), ptr(acts), ptr(out), st))

    def bwd():
        check(lib.op3_decoder_bwd(C.byref(d), N, ptr(prep), ptr(z), ptr(acts), ptr(seed), ptr(scratch), ptr(dz), 0, None, st))
    for name, fn in (("decoder fwd", fwd), ("decoder dgrad (latent)", bwd)):
        for _ in range(5):
            fn()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / 20)
        print("%s: %.4f ms per call (min of 5 x 20), median %.4f" % (name, min(ts), sorted(ts)[2]))


if __name__ == "__main__":
    main()
