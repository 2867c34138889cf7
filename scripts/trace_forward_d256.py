"""Pipeline timeline of the 128 < D <= 256 forward kernel (debug instantiation with clock64() probes in CTA (0,0)).
Usage on the GPU box:  python scripts/trace_forward_d256.py [N] [heads]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfa_b200 as mfa  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
H = int(sys.argv[2]) if len(sys.argv) > 2 else 1
D = 256
desc = mfa.AttentionDescriptor()
desc.lowPrecisionInputs = True
desc.inputPrecisionOverride = mfa.GEMMOperandPrecision.BF16
desc.matrixDimensions = (N, N, D)
desc.transposeState = (False,) * 4
desc.batchCount = H
kernel = mfa.AttentionKernel(desc.kernelDescriptor(mfa.AttentionKernelType.forward))
constants = mfa.FunctionConstantValues()
desc.setFunctionConstants(constants)
q, k, v = (torch.randn(H, N, D, device="cuda").to(torch.bfloat16) for _ in range(3))
o = torch.empty(H, N, D, device="cuda")
lse = torch.empty(H, N, device="cuda")
trace = torch.zeros(5 * 128 * 8, dtype=torch.int64, device="cuda")
lib = mfa._lib
lib.mfa_debug_forward_trace.argtypes = [ctypes.c_void_p] * 5
arr = (ctypes.c_void_p * 10)()
for slot, t in ((0, q), (1, k), (2, v), (3, o), (4, lse)):
    arr[slot] = t.data_ptr()
for _ in range(3):
    st = lib.mfa_debug_forward_trace(kernel._handle, ctypes.byref(constants._c), ctypes.byref(arr), None,
                                     ctypes.c_void_p(trace.data_ptr()))
    assert st == 0, lib.mfa_last_error()
    torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(5, 128, 8)
nb = min(N // 128, 128)
t0 = t[2, 0, 0]
print("softmax slots: 0 S ready, 1 S in regs, 2 P computed (after joint decision), 3 P store issued, 4 arrived; "
      "mma: 0 V/K ready, 1 P ready, 2 PV + next S issued; tma: 0 loop top, 1 k_empty passed, 2 v_empty passed")
for i in range(min(nb, 14)):
    print(f"i={i:2d} wg0 {(t[0, i, :5] - t0).tolist()} wg1 {(t[1, i, :5] - t0).tolist()} mma {(t[2, i, :3] - t0).tolist()} "
          f"tma {(t[4, i, :3] - t0).tolist()}")
lo, hi = 4, nb - 2
print("steady state per block:", float(np.diff(t[2, lo:hi, 1]).mean()), "cycles;",
      "softmax pass (S ready -> arrived):", float((t[0, lo:hi, 4] - t[0, lo:hi, 0]).mean()),
      "; P arrive -> MMA sees it:", float((t[2, lo:hi, 1] - np.maximum(t[0, lo:hi, 4], t[1, lo:hi, 4])).mean()),
      "; MMA issue -> next-next S ready at softmax:", float((t[0, lo + 2:hi, 0] - t[2, lo:hi - 2, 2]).mean()))
