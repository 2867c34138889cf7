"""Parity of the tensor-core backward kernels (dQ and dK/dV, TMA + tcgen05 + TMEM) with the CPU oracle, on inputs
rounded to the kernels' 16-bit memory format.  Forward runs first (it produces O and L), then backwardQuery (writes
D and dQ), then backwardKeyValue -- the reference's order (SquareAttentionTest.swift:355-368).

Stated tolerances: the reference's mixed-precision bars (D 1e-1, gradients 5e-2, RectangularAttentionTest.swift:
459-464) and, tighter, relative RMS error of every gradient <= 4e-3 for BF16 / 1.5e-3 for FP16 (P and dS are rounded
to the 16-bit MMA input type before the accumulate GEMMs: two rounded operands per gradient)."""
import numpy as np
import pytest


def _rel_rms(actual, expected):
    denom = float(np.sqrt(np.mean(expected ** 2)))
    err = float(np.sqrt(np.mean((actual - expected) ** 2)))
    return err / denom if denom > 1e-12 else err   # e.g. C == 1: dS == 0, so dK == dQ == 0 exactly


def _run(R, C, D, bf16, seed, lowMid=False, referencePolicy=False):
    import mfa_b200 as mfa
    import oracle
    from tests.attention_harness import run_attention, oracle_outputs, check

    desc = mfa.AttentionDescriptor()
    desc.lowPrecisionInputs = True
    desc.lowPrecisionIntermediates = lowMid
    desc.matrixDimensions = (R, C, D)
    desc.transposeState = (False, False, False, False)
    if referencePolicy:
        # the reference's own policy: FP16 Q, K, V and BF16 dO (AttentionDescriptor+Precisions.swift:13-23); the
        # kernels rewrite the staged dO tiles as FP16 on chip (tcgen05 kind::f16 cannot mix FP16 and BF16 operands)
        assert not bf16
        prec = desc.memoryPrecisions
        assert prec[mfa.AttentionOperand.Q] == mfa.GEMMOperandPrecision.FP16
        assert prec[mfa.AttentionOperand.dO] == mfa.GEMMOperandPrecision.BF16
    else:
        desc.inputPrecisionOverride = mfa.GEMMOperandPrecision.BF16 if bf16 else mfa.GEMMOperandPrecision.FP16
    for t in mfa.AttentionKernelType:
        assert desc.kernelDescriptor(t).backend == mfa.Backend.tcgen05, t
    net = oracle.Network(R, C, D, seed=seed, threads=8)
    prec = desc.memoryPrecisions
    net.round_inputs(int(prec[mfa.AttentionOperand.Q]), int(prec[mfa.AttentionOperand.dO]))
    out = run_attention(desc, net)
    ref = oracle_outputs(net)
    check(ref["D"], out["D"], 1e-1 if lowMid else 2e-2, "D")
    bound = 4e-3 if bf16 else 1.5e-3
    if lowMid and not bf16:
        bound = 6e-3   # L read back from FP16 (|L| ~ 8: half an ulp = 2^-8 in log2 units -> P off by up to 0.27 %)
    for name in ("dV", "dK", "dQ"):
        check(ref[name], out[name], 5e-2, name)
        rel = _rel_rms(out[name], ref[name])
        assert rel <= bound, f"{name}: relative RMS error {rel:.3e} > {bound}"
    return out, ref


SHAPES = [(128, 128, 64), (256, 256, 128), (384, 256, 64), (200, 333, 128), (77, 129, 64), (300, 17, 80),
          (129, 257, 72), (1, 1, 8), (512, 640, 96), (1024, 1024, 128)]


@pytest.mark.gpu
@pytest.mark.parametrize("R,C,D", SHAPES)
def test_backward_bf16_matches_oracle(R, C, D):
    _run(R, C, D, True, seed=R + 3 * C + D)


@pytest.mark.gpu
@pytest.mark.parametrize("R,C,D", SHAPES[:6])
def test_backward_fp16_matches_oracle(R, C, D):
    _run(R, C, D, False, seed=5 * R + C + D)


@pytest.mark.gpu
@pytest.mark.parametrize("R,C,D,bf16,referencePolicy", [(700, 900, 128, True, False), (900, 700, 64, False, True),
                                                        (2048, 2048, 128, True, False), (1000, 520, 72, False, False)])
def test_backward_traversal_split_small_grids(R, C, D, bf16, referencePolicy):
    """Few CTAs for 148 SMs: the traversal axis (keys for dQ, queries for dK/dV) is cut into ranges handled by separate
    CTAs whose partial accumulators a sum kernel adds up; ragged last ranges, ragged last blocks."""
    import mfa_b200 as mfa
    desc = mfa.AttentionDescriptor()
    desc.lowPrecisionInputs = True
    if not referencePolicy:
        desc.inputPrecisionOverride = mfa.GEMMOperandPrecision.BF16 if bf16 else mfa.GEMMOperandPrecision.FP16
    desc.matrixDimensions = (R, C, D)
    desc.transposeState = (False, False, False, False)
    constants = mfa.FunctionConstantValues()
    desc.setFunctionConstants(constants)
    for t in (mfa.AttentionKernelType.backwardQuery, mfa.AttentionKernelType.backwardKeyValue):
        assert mfa.AttentionKernel(desc.kernelDescriptor(t)).launchCount(constants) == 2, t
    _run(R, C, D, bf16, seed=R + C + D, referencePolicy=referencePolicy)


@pytest.mark.gpu
def test_backward_low_precision_intermediates_bf16():
    """L stored FP16, D stored BF16 (AttentionDescriptor+Precisions.swift:81-87) and read back by dK/dV."""
    _run(256, 384, 64, True, seed=4, lowMid=True)


@pytest.mark.gpu
@pytest.mark.parametrize("R,C,D", [(128, 128, 64), (256, 256, 128), (200, 333, 128), (77, 129, 64), (300, 17, 80),
                                   (129, 257, 72), (1, 1, 8), (512, 640, 96), (640, 1024, 128)])
@pytest.mark.parametrize("lowMid", [False, True])
def test_backward_reference_policy_fp16_inputs_bf16_dO(R, C, D, lowMid):
    """The reference's unmodified low-precision descriptor (FP16 Q/K/V, BF16 dO; with lowPrecisionIntermediates also
    FP16 L and BF16 D) on the tensor-core kernels."""
    _run(R, C, D, False, seed=7 * R + C + D, lowMid=lowMid, referencePolicy=True)


@pytest.mark.gpu
def test_config3_fwd_bwd_n2048_d64():
    """BASELINE.json configs[2]: forward + backward (dQ, dK/dV) N=2048 D=64 on one B200 through the tensor-core family:
    the reference's policy (FP16 Q/K/V + BF16 dO, as the reference would run "fp16"), all-FP16 and all-BF16."""
    _run(2048, 2048, 64, False, seed=2, referencePolicy=True)
    _run(2048, 2048, 64, False, seed=0)
    _run(2048, 2048, 64, True, seed=1)
