"""Pipeline timeline of the tcgen05 forward kernel (debug instantiation with clock64() probes).
Usage on the GPU box:  python scripts/trace_forward.py [R] [heads]  -> prints per-iteration cycle deltas."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfa_b200 as mfa  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1
C, D = 4096, 128

desc = mfa.AttentionDescriptor()
desc.lowPrecisionInputs = True
desc.inputPrecisionOverride = mfa.GEMMOperandPrecision.BF16
desc.matrixDimensions = (R, C, D)
desc.transposeState = (False,) * 4
desc.batchCount = H
kernel = mfa.AttentionKernel(desc.kernelDescriptor(mfa.AttentionKernelType.forward))
constants = mfa.FunctionConstantValues()
desc.setFunctionConstants(constants)

q = torch.randn(H, R, D, device="cuda").to(torch.bfloat16)
k = torch.randn(H, C, D, device="cuda").to(torch.bfloat16)
v = torch.randn(H, C, D, device="cuda").to(torch.bfloat16)
o = torch.empty(H, R, D, device="cuda")
lse = torch.empty(H, R, device="cuda")
trace = torch.zeros(4 * 128 * 8, dtype=torch.int64, device="cuda")

lib = mfa._lib
lib.mfa_debug_forward_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                        ctypes.c_void_p]
arr = (ctypes.c_void_p * 10)()
for slot, t in ((0, q), (1, k), (2, v), (3, o), (4, lse)):
    arr[slot] = t.data_ptr()
for _ in range(3):
    st = lib.mfa_debug_forward_trace(kernel._handle, ctypes.byref(constants._c), ctypes.byref(arr), None,
                                     ctypes.c_void_p(trace.data_ptr()))
    assert st == 0, lib.mfa_last_error()
    torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(4, 128, 8)
nb = C // 128
t0 = t[2, 0, 0]
print(f"R={R} H={H}: 128-key blocks; softmax slots: 0 step start, 1 S(i+1) waited + ld issued, 2 max(i+1) done, 3 exp(i) done, "
      "4 P arrived, 5 pair exchange + max update done; mma slots: 0 V ready, 1 P ready, 2 PV issued, 3 K ready, 4 S(i+3) issued")
for j in range(4, 12):
    a = (t[0, j, :6] - t0).tolist()
    b = (t[1, j, :6] - t0).tolist()
    m = (t[2, j, :5] - t0).tolist()
    print(f"i={j:2d} lo {a}  hi {b}  mma {m}")
lo, hi = 4, nb - 4
per_iter = np.diff(t[2, lo:hi, 1]).mean()
print(f"steady-state period per 128-key block: {per_iter:.0f} cycles (floors: ~1000 tensor, 1024 MUFU)")
for role in (0, 1):
    sm = t[role, lo:hi]
    print(f"softmax half{role}: wait S(i+1)+ld issue {np.mean(sm[:,1]-sm[:,0]):.0f}, exp chunk0 + max(i+1) {np.mean(sm[:,2]-sm[:,1]):.0f}, "
          f"exp chunk1 {np.mean(sm[:,3]-sm[:,2]):.0f}, wait_st+arrive {np.mean(sm[:,4]-sm[:,3]):.0f}, "
          f"exchange+update {np.mean(sm[:,5]-sm[:,4]):.0f}, loop overhead {np.mean(sm[1:,0]-sm[:-1,5]):.0f}")
mm = t[2, lo:hi]
print(f"mma: wait P {np.mean(mm[:,1]-mm[:,0]):.0f}; issue PV {np.mean(mm[:,2]-mm[:,1]):.0f}; wait K {np.mean(mm[:,3]-mm[:,2]):.0f}; "
      f"issue S {np.mean(mm[:,4]-mm[:,3]):.0f}; V wait {np.mean(mm[1:,0]-mm[:-1,4]):.0f}")
