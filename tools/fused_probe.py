"""Times op3_dynamics_fwd / bwd (fused chains vs layer-by-layer launches) for several batch sizes: CUDA events, warm L2."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from op3_b200 import _lib
    from op3_b200._lib import check, ptr
    from gpu_util import build_model
    lib = _lib.load()
    K, A = int(os.environ.get("K", 4)), int(os.environ.get("A", 4))
    m, _ = build_model(A, K, seed=1)
    eng = m.engine
    st = torch.cuda.current_stream().cuda_stream
    for B in [int(x) for x in (sys.argv[1:] or ["1", "8", "64", "148", "592", "4096"])]:
        N, R = B * K, 128
        z = torch.randn(N, R, device="cuda") * 0.7
        ac = torch.randn(B, 4, device="cuda") if A else None
        d = eng.dims(B)
        ps = eng.params_struct()
        acts = torch.zeros(lib.op3_dynamics_acts_floats(C.byref(d)), device="cuda")
        oz, o1, o2 = torch.zeros(N, R, device="cuda"), torch.empty(N, 64, device="cuda"), torch.empty(N, 64, device="cuda")
        gflat = torch.zeros_like(eng.flat_params())
        gs = eng._struct_for(gflat)
        dz = torch.empty(N, R, device="cuda")
        scratch = torch.empty(lib.op3_dynamics_scratch_floats(C.byref(d)), device="cuda")
        row = ["B=%5d" % B]
        for fused in (1, 0):
            lib.op3_set_fused_mlp(fused)
            for what in ("fwd", "bwd"):
                def call():
                    if what == "fwd":
                        check(lib.op3_dynamics_fwd(C.byref(d), C.byref(ps), ptr(z), ptr(ac), ptr(acts), ptr(oz), ptr(o1), ptr(o2), st))
                    else:
                        check(lib.op3_dynamics_bwd(C.byref(d), C.byref(ps), ptr(z), ptr(ac), ptr(acts), ptr(oz), ptr(o1), ptr(o2),
                                                   ptr(dz), C.byref(gs), ptr(scratch), st))
                for _ in range(3):
                    call()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(10):
                    call()
                e1.record()
                torch.cuda.synchronize()
                row.append("%s %s %8.1f us" % ("fused" if fused else "layer", what, e0.elapsed_time(e1) * 100))
        lib.op3_set_fused_mlp(1)
        print("  ".join(row), flush=True)


if __name__ == "__main__":
    main()
