"""Runs forward + dQ + dK/dV once per step at N=4096, D=128, bf16, 32 heads (for ncu captures of the backward kernels)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.bench_configs import run
import mfa_b200 as mfa
print(run(4096, 128, mfa.GEMMOperandPrecision.BF16, 32, steps=2))
