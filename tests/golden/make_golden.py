"""Generates tests/golden/*.npz from the CPU oracle (oracle/network_oracle.c, the C restatement of the
reference's `Network`).  The reference itself ships no golden vectors and cannot run in this image (Swift +
Metal), so these fixtures pin OUR oracle build against itself across machines/compilers, and give the GPU
tests vectors that do not depend on rebuilding the oracle.  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [  # (name, R, C, D, seed, input rounding)
    ("fp32_r10_c10_d3", 10, 10, 3, 1, None),            # SquareAttentionTest.swift:6
    ("fp32_r25_c25_d2", 25, 25, 2, 2, None),            # :12
    ("fp32_r64_c64_d40", 64, 64, 40, 3, None),          # :20
    ("fp32_r93_c77_d32", 93, 77, 32, 4, None),          # rectangular
    ("fp32_r128_c128_d64", 128, 128, 64, 5, None),      # BASELINE.json configs[0]
    ("bf16_r256_c384_d128", 256, 384, 128, 6, oracle.BF16),
    ("fp16_r200_c136_d64", 200, 136, 64, 7, oracle.FP16),
]

for name, R, C, D, seed, rounding in CASES:
    net = oracle.Network(R, C, D, seed=seed)
    if rounding is not None:
        net.round_inputs(rounding)
    O, L = net.inferenceAttention(with_L=True)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), Q=net.Q, K=net.K, V=net.V, dO=net.dO, O=O, L=L,
                        D=net.createDTerms(), dV=net.derivativeV(), dK=net.derivativeK(), dQ=net.derivativeQ(),
                        meta=np.array([R, C, D, seed, -1 if rounding is None else rounding]))
    print("wrote", name)
