"""Pins the CPU oracle (oracle/network_oracle.c = C restatement of the reference's `Network`,
Tests/FlashAttentionTests/Utilities/Network.swift:70-403).

The reference ships no golden vectors for this path and cannot run here (Swift + Metal), so the oracle is
"parity unpinned" against reference-generated data; what pins it is:
  1. an independent float64 matrix-form implementation (oracle/oracle_np.py) on the reference's own 20 test
     shapes (SquareAttentionTest.swift:6-25);
  2. central finite differences of the reference's loss Phi = sum dO*O (Network.swift:314-326), the method of
     Documentation/Archive/FiniteDifferencingTest.swift:85-134;
  3. softmax identities;
  4. committed golden fixtures (tests/golden/*.npz) that freeze today's outputs across machines."""
import glob
import os

import numpy as np
import pytest

import oracle
from oracle.oracle_np import attention_f64

REFERENCE_SHAPES = [
    (10, 3), (10, 80), (8, 2), (9, 2), (23, 2), (24, 2), (25, 2), (192, 77), (192, 80), (93, 32),
    (99, 35), (64, 32), (64, 34), (64, 36), (64, 40), (32, 64), (4, 1), (4, 2), (384, 95), (777, 199),
]


@pytest.mark.parametrize("N,D", REFERENCE_SHAPES)
def test_oracle_matches_float64_formulation(N, D):
    net = oracle.Network(N, N, D, seed=N * 31 + D, threads=4)
    ref = attention_f64(net.Q, net.K, net.V, net.dO)
    O, L = net.inferenceAttention(with_L=True)
    got = dict(O=O, L=L, D=net.createDTerms(), dV=net.derivativeV(), dK=net.derivativeK(), dQ=net.derivativeQ())
    for name, value in got.items():
        scale = max(1.0, float(np.abs(ref[name]).max()))
        # FP32 sequential accumulation over N (and N*D) terms vs float64
        assert np.abs(value - ref[name]).max() <= 2e-5 * scale, name


def test_rectangular_and_threaded_variants_agree():
    net = oracle.Network(57, 131, 24, seed=9)
    single = dict(O=net.inferenceAttention(), dV=net.derivativeV(), dK=net.derivativeK(), dQ=net.derivativeQ(),
                  D=net.createDTerms())
    net.threads = 4
    multi = dict(O=net.inferenceAttention(), dV=net.derivativeV(), dK=net.derivativeK(), dQ=net.derivativeQ(),
                 D=net.createDTerms())
    for name in single:
        # per-row arithmetic is identical; only dV/dK (sums over rows) change their summation order
        tol = 0 if name in ("O", "dQ", "D") else 2e-6
        assert np.abs(single[name] - multi[name]).max() <= tol, name


def test_gradients_match_finite_differences():
    """Documentation/Archive/FiniteDifferencingTest.swift:85-134, asserted instead of printed."""
    net = oracle.Network(6, 7, 4, seed=5)
    analytic = {"Q": net.derivativeQ(), "K": net.derivativeK(), "V": net.derivativeV()}
    step = 1e-2
    rng = np.random.default_rng(0)
    for name in ("Q", "K", "V"):
        tensor = getattr(net, name)
        for _ in range(6):
            idx = tuple(rng.integers(0, s) for s in tensor.shape)
            original = tensor[idx]
            tensor[idx] = original + step
            plus = net.loss()
            tensor[idx] = original - step
            minus = net.loss()
            tensor[idx] = original
            numeric = (plus - minus) / (2 * step)
            assert abs(numeric - analytic[name][idx]) <= 2e-3 * max(1.0, abs(numeric)), (name, idx)


def test_softmax_identities():
    net = oracle.Network(33, 45, 16, seed=3)
    O, L = net.inferenceAttention(with_L=True)
    ones = oracle.Network(33, 45, 16, seed=3)
    ones.V = np.ones_like(ones.V)
    assert np.abs(ones.inferenceAttention() - 1.0).max() < 5e-6          # rows of P sum to 1
    S = (net.Q.astype(np.float64) @ net.K.astype(np.float64).T) / np.sqrt(16)
    lse = np.log(np.exp(S - S.max(1, keepdims=True)).sum(1)) + S.max(1)
    assert np.abs(L - lse).max() < 1e-5                                     # L is the natural-log LSE
    assert np.abs(net.createDTerms() - (net.dO * O).sum(1)).max() < 1e-5    # D = rowsum(dO * O)


def test_box_muller_inputs_are_standard_normal_and_seeded():
    a, b = oracle.Network(64, 64, 64, seed=1), oracle.Network(64, 64, 64, seed=1)
    c = oracle.Network(64, 64, 64, seed=2)
    assert np.array_equal(a.Q, b.Q) and not np.array_equal(a.Q, c.Q)
    for t in (a.Q, a.K, a.V, a.dO):
        assert abs(t.mean()) < 0.06 and abs(t.std() - 1) < 0.06


def test_encode_decode_follow_reference_buffer_rules():
    """MTLContext+Buffers.swift:31-44: FP16 = round to nearest even, BF16 = truncation."""
    x = np.array([1.0, 1.0 + 2.0 ** -9, 1.0 + 3 * 2.0 ** -9, -2.5, 65504.0, 1e-8, 3.1415927], np.float32)
    assert np.array_equal(oracle.roundtrip(x, oracle.FP16), x.astype(np.float16).astype(np.float32))
    bits = x.view(np.uint32) & np.uint32(0xFFFF0000)
    assert np.array_equal(oracle.roundtrip(x, oracle.BF16), bits.view(np.float32))
    assert np.array_equal(oracle.roundtrip(x, oracle.FP32), x)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))))
def test_oracle_reproduces_golden_fixtures(path):
    g = np.load(path)
    R, C, D, seed, rounding = (int(v) for v in g["meta"])
    net = oracle.Network(R, C, D, seed=seed)
    if rounding >= 0:
        net.round_inputs(rounding)
    for name in ("Q", "K", "V", "dO"):
        assert np.array_equal(getattr(net, name), g[name]), name   # seeded generator is bit-stable
    O, L = net.inferenceAttention(with_L=True)
    got = dict(O=O, L=L, D=net.createDTerms(), dV=net.derivativeV(), dK=net.derivativeK(), dQ=net.derivativeQ())
    for name, value in got.items():
        # same C source, same flags (-ffp-contract=off): only libm's expf/logf may differ by an ulp across hosts
        assert np.abs(value - g[name]).max() <= 2e-6 * max(1.0, float(np.abs(g[name]).max())), name
