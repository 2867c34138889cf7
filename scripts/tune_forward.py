"""Sweep the forward kernel's warpgroup stagger on the GPU box and print TFLOP/s for each setting."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfa_b200 as mfa

H, N, D = 64, 4096, 128
desc = mfa.AttentionDescriptor()
desc.lowPrecisionInputs = True
desc.inputPrecisionOverride = mfa.GEMMOperandPrecision.BF16
desc.matrixDimensions = (N, N, D)
desc.transposeState = (False,) * 4
desc.batchCount = H
kernel = mfa.AttentionKernel(desc.kernelDescriptor(mfa.AttentionKernelType.forward))
c = mfa.FunctionConstantValues(); desc.setFunctionConstants(c)
Op = mfa.AttentionOperand
bufs = {Op.Q: torch.randn(H, N, D, device="cuda").bfloat16(), Op.K: torch.randn(H, N, D, device="cuda").bfloat16(),
        Op.V: torch.randn(H, N, D, device="cuda").bfloat16(), Op.O: torch.empty(H, N, D, device="cuda"),
        Op.L: torch.empty(H, N, device="cuda")}
ptrs = {op: t.data_ptr() for op, t in bufs.items()}
stream = torch.cuda.current_stream().cuda_stream
values = [int(x) for x in sys.argv[1:]] or [0, 200, 400, 500, 600, 700, 800, 1000]
for v in values:
    mfa._lib.mfa_debug_set_forward_stagger(ctypes.c_uint32(v))
    for _ in range(5):
        kernel.encode(c, ptrs, stream)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(30):
        kernel.encode(c, ptrs, stream)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 30
    print(f"stagger {v:5d}: {ms:.4f} ms  {4*N*N*D*H/ms/1e9:.1f} TFLOP/s", flush=True)
