"""Debug: timeline of the tcgen05 conv issuer (needs a build with OP3_NVCC_EXTRA=-DOP3_TC_TRACE).
Runs the decoder forward (three 32->32 tensor-core layers per call) at N = 256 slot images and prints where the
MMA-issuing thread spent its cycles."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import op3_b200
    from op3_b200 import _lib
    from op3_b200._lib import check, ptr
    from op3_b200.torch.op3_modules.op3_model import create_model_v2
    lib = _lib.load()
    op3_b200.set_precision(sys.argv[1] if len(sys.argv) > 1 else "f16x3")
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    variant = dict(refinement_model_type="size_dependent_conv", decoder_model_type="reg", dynamics_model_type="reg_ac32",
                   sto_repsize=64, det_repsize=64, K=4, extra_args=dict(beta=1e-2, deterministic_sampling=False))
    model = create_model_v2(variant, 64, 64, 4).to(dev)
    eng = model.engine
    B, K = 64, 4
    N = B * K
    d = eng.dims(B)
    st = torch.cuda.current_stream(dev).cuda_stream
    prep = eng.prepared(st)
    z = torch.randn(N, 128, device=dev)
    acts = torch.empty(lib.op3_decoder_acts_floats(C.byref(d), N), device=dev)
    out = torch.empty(N, 64 * 64, 4, device=dev)
    f = lib.op3_debug_tc_trace
    f.argtypes = [C.c_void_p, C.c_int]
    if os.environ.get("TCQ_DBG"):
        fl = lib.op3_debug_tcq_flags
        fl.argtypes = [C.c_int]
        fl(int(os.environ["TCQ_DBG"]))
    if os.environ.get("TCQ_PAIR"):
        lib.op3_set_tcq_pair(int(os.environ["TCQ_PAIR"]))
    for _ in range(3):
        check(lib.op3_decoder_fwd(C.byref(d), N, ptr(prep), ptr(z), ptr(acts), ptr(out), st), "decoder_fwd")
    torch.cuda.synchronize()
    f(None, 1)
    g = lib.op3_debug_tcq_epi
    g.argtypes = [C.c_void_p, C.c_int]
    g(None, 1)
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        check(lib.op3_decoder_fwd(C.byref(d), N, ptr(prep), ptr(z), ptr(acts), ptr(out), st), "decoder_fwd")
    e1.record()
    torch.cuda.synchronize()
    buf = np.zeros((256, 4), dtype=np.uint64)
    f(buf.ctypes.data, 0)
    t = buf[:148].astype(np.float64)
    launches = reps * (4 if (len(sys.argv) < 2 or sys.argv[1] == "f16x3") else 3)   # f16x3: the output layer runs on the tensor kernel too
    print("decoder fwd %.3f ms per call" % (e0.elapsed_time(e1) / reps))
    print("per launch, mean over CTAs: wait operands %.0f cyc, wait accumulators %.0f cyc, total %.0f cyc, MMAs %.0f"
          % tuple(t.mean(0) / launches))
    tot, nm = t[:, 2].mean() / launches, t[:, 3].mean() / launches
    print("cycles per MMA pair while issuing: %.1f" % ((tot - t[:, 0].mean() / launches - t[:, 1].mean() / launches) / (nm / 2)))
    e = np.zeros(8, dtype=np.uint64)
    g(e.ctypes.data, 0)
    if e[3]:
        n = float(e[3])
        print("quint epilogue (warp 0 of every CTA, cycles per tile): wait for accumulators %.0f | TMEM loads + shuffles (-> slot released) %.0f | "
              "group barrier %.0f | boundary fix-up %.0f | scale + bias + ELU + stores %.0f | second barrier %.0f"
              % (e[0] / n, e[1] / n, e[2] / n, e[4] / n, e[5] / n, e[6] / n))
    print("max over CTAs total %.0f, min %.0f" % (t[:, 2].max() / launches, t[:, 2].min() / launches))


if __name__ == "__main__":
    main()
