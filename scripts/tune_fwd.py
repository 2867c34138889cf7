"""Times the forward at (4096,128) and (4096,64), 64 heads, with the library named by MFA_B200_LIBRARY."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfa_b200 as mfa
KT, Op, P = mfa.AttentionKernelType, mfa.AttentionOperand, mfa.GEMMOperandPrecision
res = {"lib": os.path.basename(mfa.library_path())}
for N, D, H in ((4096, 128, 64), (4096, 64, 64)):
    desc = mfa.AttentionDescriptor(); desc.lowPrecisionInputs = True; desc.inputPrecisionOverride = P.BF16
    desc.matrixDimensions = (N, N, D); desc.transposeState = (False,) * 4; desc.batchCount = H
    bufs = {Op.Q: torch.randn(H, N, D, device="cuda").bfloat16(), Op.K: torch.randn(H, N, D, device="cuda").bfloat16(),
            Op.V: torch.randn(H, N, D, device="cuda").bfloat16(), Op.O: torch.empty(H, N, D, device="cuda"),
            Op.L: torch.empty(H, N, device="cuda")}
    ptrs = {op: t.data_ptr() for op, t in bufs.items()}
    c = mfa.FunctionConstantValues(); desc.setFunctionConstants(c)
    k = mfa.AttentionKernel(desc.kernelDescriptor(KT.forward))
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(5): k.encode(c, ptrs, st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(40): k.encode(c, ptrs, st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 40
    res[f"D{D}"] = round(4 * N * N * D * H / ms / 1e9, 1)
print(json.dumps(res), flush=True)
