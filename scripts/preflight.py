"""CPU-side preflight run before every gpurun: the library loads, every descriptor the GPU tests use maps to a
kernel object (no device needed for creation), and the Python sources parse."""
import ast
import glob
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mfa_b200 as mfa  # noqa: E402

for f in glob.glob("tests/*.py") + glob.glob("scripts/*.py") + ["bench.py", "__graft_entry__.py"]:
    ast.parse(open(f).read(), f)
n = 0
for low_in in (False, True):
    for low_mid in (False, True):
        for bf16 in (False, True):
            for D in (1, 8, 16, 64, 72, 80, 96, 128, 136, 256, 300, 512):
                for tr in ((False,) * 4, (True, False, True, False)):
                    d = mfa.AttentionDescriptor()
                    d.lowPrecisionInputs, d.lowPrecisionIntermediates = low_in, low_mid
                    d.matrixDimensions, d.transposeState = (300, 200, D), tr
                    if bf16:
                        d.inputPrecisionOverride = mfa.GEMMOperandPrecision.BF16
                    for t in mfa.AttentionKernelType:
                        k = mfa.AttentionKernel(d.kernelDescriptor(t))
                        assert k.threadgroupSize > 0 and k.blockDimensions[0] > 0
                        n += 1
print(f"preflight ok: {n} kernel objects created; {mfa.version()}")
