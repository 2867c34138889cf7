"""Device-timed throughput of the BASELINE.json configs other than the headline one (which bench.py owns):
forward + dQ + dK/dV at N=2048, D=64 (configs[2]) and a few neighbours, in the reference's units
(GINSTRS: (2D+5) N^2 fwd, (3D+5) N^2 dQ, (4D+5) N^2 dK/dV -- README.md:108-124) and GEMM TFLOP/s.
Usage (GPU box):  python scripts/bench_configs.py [heads]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfa_b200 as mfa
from bench import ClockSampler   # NVML clock / throttle-reason sampler shared with bench.py

KT, Op, P = mfa.AttentionKernelType, mfa.AttentionOperand, mfa.GEMMOperandPrecision


def run(N, D, precision, H, steps=20, transpose=(False,) * 4):
    desc = mfa.AttentionDescriptor()
    desc.lowPrecisionInputs = True
    desc.inputPrecisionOverride = precision   # None = the reference's policy: FP16 Q/K/V, BF16 dO
    desc.matrixDimensions = (N, N, D)
    desc.transposeState = tuple(transpose)   # (timing only: the buffers' contents are random either way)
    desc.batchCount = H
    dt = torch.bfloat16 if precision == P.BF16 else torch.float16
    dt_dO = torch.bfloat16 if precision is None else dt
    bufs = {Op.Q: torch.randn(H, N, D, device="cuda").to(dt), Op.K: torch.randn(H, N, D, device="cuda").to(dt),
            Op.V: torch.randn(H, N, D, device="cuda").to(dt), Op.dO: torch.randn(H, N, D, device="cuda").to(dt_dO),
            Op.O: torch.empty(H, N, D, device="cuda"), Op.L: torch.empty(H, N, device="cuda"),
            Op.D: torch.empty(H, N, device="cuda"), Op.dQ: torch.empty(H, N, D, device="cuda"),
            Op.dK: torch.empty(H, N, D, device="cuda"), Op.dV: torch.empty(H, N, D, device="cuda")}
    ptrs = {op: t.data_ptr() for op, t in bufs.items()}
    c = mfa.FunctionConstantValues()
    desc.setFunctionConstants(c)
    stream = torch.cuda.current_stream().cuda_stream
    out = {"N": N, "D": D, "dtype": precision.name if precision is not None else "FP16 Q/K/V + BF16 dO (reference policy)",
           "heads": H}
    work = {KT.forward: (2 * D + 5, 4), KT.backwardQuery: (3 * D + 5, 6), KT.backwardKeyValue: (4 * D + 5, 8)}
    for t in KT:
        kd = desc.kernelDescriptor(t)
        k = mfa.AttentionKernel(kd)
        for _ in range(3):
            k.encode(c, ptrs, stream)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(torch.cuda.current_device()) as sampler:
            a.record()
            for _ in range(steps):
                k.encode(c, ptrs, stream)
            b.record()
            torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
        fma, gemm = work[t]
        out[t.name] = {"ms": round(ms, 4), "ginstrs": round(fma * N * N * H / ms / 1e6, 1),
                       "tflops": round(gemm * N * N * D * H / ms / 1e9, 1), "kernel": k.sourceName(),
                       "clocks": sampler.summary()}
    return out


if __name__ == "__main__":
    H = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    for N, D, prec, heads in ((8192, 256, P.BF16, 16), (8192, 256, P.BF16, 1), (2048, 64, None, H), (2048, 64, P.FP16, H),
                              (2048, 64, P.BF16, H), (4096, 128, None, 64), (4096, 128, P.BF16, 64), (4096, 64, P.BF16, 64),
                              (2048, 64, P.FP16, 1), (4096, 128, P.BF16, 1), (4096, 256, P.BF16, 16), (4096, 256, None, 16),
                              (4096, 192, P.BF16, 16)):
        print(json.dumps(run(N, D, prec, heads)), flush=True)
    # transposed operands (the layout-generic kernels); timing only
    for N, D, heads, tr in ((4096, 128, 32, (True,) * 4), (4096, 128, 32, (False, True, False, False)),
                            (2048, 64, 64, (True,) * 4), (4096, 256, 16, (True,) * 4)):
        r = run(N, D, P.BF16, heads, transpose=tr)
        r["transposeState(Q,K,V,O)"] = list(tr)
        print(json.dumps(r), flush=True)
