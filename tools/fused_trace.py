"""Cycle breakdown of the fused dynamics forward kernel (build with OP3_NVCC_EXTRA=-DOP3_FUSED_TRACE)."""
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
    raw = C.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else lib
    fn = raw.op3_debug_fused_trace
    fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    K, A = 4, 4
    m, _ = build_model(A, K, seed=1)
    eng = m.engine
    st = torch.cuda.current_stream().cuda_stream
    for B in (1, 64):
        N, R = B * K, 128
        z = torch.randn(N, R, device="cuda") * 0.7
        ac = torch.randn(B, 4, device="cuda")
        d = eng.dims(B)
        ps = eng.params_struct()
        acts = torch.zeros(lib.op3_dynamics_acts_floats(C.byref(d)), device="cuda")
        oz, o1, o2 = torch.zeros(N, R, device="cuda"), torch.empty(N, 64, device="cuda"), torch.empty(N, 64, device="cuda")
        for _ in range(3):
            check(lib.op3_dynamics_fwd(C.byref(d), C.byref(ps), ptr(z), ptr(ac), ptr(acts), ptr(oz), ptr(o1), ptr(o2), st))
        torch.cuda.synchronize()
        buf = (C.c_ulonglong * 8)()
        fn(None, 1)
        check(lib.op3_dynamics_fwd(C.byref(d), C.byref(ps), ptr(z), ptr(ac), ptr(acts), ptr(oz), ptr(o1), ptr(o2), st))
        torch.cuda.synchronize()
        fn(buf, 0)
        print("B=%d kernel %d cyc | in cta_gemm %d | tile waits %d | epilogues %d | tiles %d | calls %d | issue %d | compute %d"
              % ((B,) + tuple(buf[:8])))


if __name__ == "__main__":
    main()
