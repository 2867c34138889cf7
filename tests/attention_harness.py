"""GPU test harness: what the reference's tests do around the kernels
(Tests/FlashAttentionTests/Attention/SquareAttentionTest.swift:214-555 and
RectangularAttentionTest.swift:39-473), against the C ABI instead of Metal.

  * inputs come from the seeded CPU oracle `Network`, are transposed on request
    (Rectangular:88-138) and encoded to each operand's memory precision
    (MTLContext+Buffers.swift:5-45: FP16 = RNE, BF16 = truncate);
  * every device buffer gets an equal-length random tail in [-20, 20] to expose out-of-bounds
    accesses (MTLContext+Buffers.swift:9-18), outputs are NaN-poisoned (Square:286);
  * kernels run in the reference's order fwd -> dQ -> dK/dV (Square:355-368);
  * results are decoded, un-transposed and converted back to the oracle's units:
    L / log2(e), D * sqrt(D) (Square:405-413).
"""
import os

import numpy as np
import torch

import mfa_b200 as mfa
from mfa_b200 import AttentionKernelType as KT
from mfa_b200 import AttentionOperand as Op
import oracle

LOG2E = 1.44269504089

_SEQ = {Op.Q: "R", Op.O: "R", Op.dO: "R", Op.dQ: "R", Op.K: "C", Op.V: "C", Op.dK: "C", Op.dV: "C"}


def _transpose_in(a):  # [..., seq, D] -> [..., D, seq]
    return np.ascontiguousarray(np.swapaxes(a, -1, -2))


def _device_buffer(raw: np.ndarray, rng, precision, pad=True):
    """raw: uint16 or float32 memory image. Appends the OOB-detection tail and uploads."""
    flat = raw.reshape(-1)
    if pad:
        tail = oracle.encode(rng.uniform(-20, 20, size=flat.size).astype(np.float32), precision)
        flat = np.concatenate([flat, tail.reshape(-1)])
    if flat.dtype == np.uint16:
        return torch.from_numpy(flat.view(np.int16)).cuda()
    return torch.from_numpy(flat).cuda()


def run_attention(desc: "mfa.AttentionDescriptor", network, types=(KT.forward, KT.backwardQuery, KT.backwardKeyValue),
                  inputs=None, stream=None, return_raw=False):
    """Runs the requested kernels for `desc` on the current CUDA device. `network` supplies Q,K,V,dO as
    float32 [seq, D] (or [batch, seq, D] when desc.batchCount > 1). Returns {name: float32 array} in the
    oracle's layout/units."""
    R, C, D = desc.matrixDimensions
    batch = max(1, desc.batchCount)
    tQ, tK, tV, tO = desc.transposeState
    transposed = {Op.Q: tQ, Op.K: tK, Op.V: tV, Op.O: tO, Op.dO: tO, Op.dV: tV, Op.dK: tK, Op.dQ: tQ}
    precisions = desc.memoryPrecisions
    rng = np.random.default_rng(12345)

    host_in = inputs or {Op.Q: network.Q, Op.K: network.K, Op.V: network.V, Op.dO: network.dO}
    dev = {}
    for op, arr in host_in.items():
        a = np.asarray(arr, np.float32)
        if transposed[op]:
            a = _transpose_in(a)
        dev[op] = _device_buffer(oracle.encode(a, int(precisions[op])), rng, int(precisions[op]))

    def out_buffer(op, count):
        prec = precisions[op]
        if prec == mfa.GEMMOperandPrecision.FP32:
            buf = torch.full((2 * count,), float("nan"), dtype=torch.float32, device="cuda")
        else:
            buf = torch.full((2 * count,), -1, dtype=torch.int16, device="cuda")  # 0xFFFF = NaN in both 16-bit formats
        return buf

    counts = {Op.O: batch * R * D, Op.L: batch * R, Op.D: batch * R, Op.dV: batch * C * D, Op.dK: batch * C * D,
              Op.dQ: batch * R * D}
    for op, n in counts.items():
        dev[op] = out_buffer(op, n)

    constants = mfa.FunctionConstantValues()
    desc.setFunctionConstants(constants)
    kernels = {}
    for t in types:
        kernels[t] = mfa.AttentionKernel(desc.kernelDescriptor(t))
    buffers = {op: t.data_ptr() for op, t in dev.items()}
    cu_stream = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
    for t in (KT.forward, KT.backwardQuery, KT.backwardKeyValue):
        if t in kernels:
            kernels[t].encode(constants, buffers, cu_stream)
    torch.cuda.synchronize()

    produced = []
    if KT.forward in kernels:
        produced += [Op.O, Op.L]
    if KT.backwardQuery in kernels:
        produced += [Op.D, Op.dQ]
    if KT.backwardKeyValue in kernels:
        produced += [Op.dV, Op.dK]

    out, tails_ok = {}, True
    for op in produced:
        n = counts[op]
        t = dev[op].cpu().numpy()
        raw, tail = t[:n], t[n:]
        # the poisoned tail must be untouched (out-of-bounds write detection)
        if t.dtype == np.float32:
            tails_ok &= bool(np.isnan(tail).all())
            vals = raw.copy()
        else:
            tails_ok &= bool((tail == -1).all())
            vals = oracle.decode(raw.view(np.uint16), int(precisions[op]))
        if op in (Op.L, Op.D):
            vals = vals.reshape((batch, R) if batch > 1 else (R,))
        else:
            seq = R if _SEQ[op] == "R" else C
            if transposed[op]:
                vals = np.swapaxes(vals.reshape((batch, D, seq) if batch > 1 else (D, seq)), -1, -2)
            else:
                vals = vals.reshape((batch, seq, D) if batch > 1 else (seq, D))
        out[op.name] = np.ascontiguousarray(vals, np.float32)
    assert tails_ok, "a kernel wrote past the end of an output buffer"
    if return_raw:
        return out
    if "L" in out:
        out["L"] = out["L"] / np.float32(LOG2E)            # stored in log2 units (Square:408-410)
    if "D" in out:
        out["D"] = out["D"] * np.float32(np.sqrt(D))        # stored pre-scaled by 1/sqrt(D) (Square:411-413)
    return out


def oracle_outputs(network, backward=True):
    O, L = network.inferenceAttention(with_L=True)
    out = {"O": O, "L": L}
    if backward:
        out.update(D=network.createDTerms(), dV=network.derivativeV(), dK=network.derivativeK(),
                   dQ=network.derivativeQ())
    return out


# Measured errors of every check() in the session; tests/conftest.py writes them to gpurun_out/parity_tests.jsonl so
# that the achieved accuracy (not just pass / fail) of each GPU test is on record.
RECORDS = []


def record(name, expected, actual, tolerance=None):
    e, a = np.asarray(expected, np.float64), np.asarray(actual, np.float64)
    if e.size == 0:
        return
    err = np.abs(e - a)
    rms = float(np.sqrt(np.mean(e * e)))
    floor = max(1e-3 * rms, 1e-30)
    with np.errstate(invalid="ignore", divide="ignore"):
        RECORDS.append({"test": os.environ.get("PYTEST_CURRENT_TEST", "").split(" (")[0], "output": name, "shape": list(e.shape), "tolerance": tolerance,
                        "max_abs": float(np.nanmax(err)),
                        "max_rel": float(np.nanmax(err / np.maximum(np.abs(e), floor))),
                        "rel_rms": float(np.sqrt(np.nanmean(err * err)) / max(rms, 1e-30))})


def check(expected, actual, tolerance, name=""):
    """check() of SquareAttentionTest.swift:513-536, but asserting (the reference only prints)."""
    expected, actual = np.asarray(expected), np.asarray(actual)
    assert expected.shape == actual.shape, (name, expected.shape, actual.shape)
    record(name, expected, actual, tolerance)
    err = np.abs(expected - actual)
    bad = (err > tolerance) | np.isnan(err)
    # NaN/Inf-vs-NaN/Inf pairs are skipped by the reference (Square:521-524)
    skip = (~np.isfinite(expected)) & (~np.isfinite(actual))
    bad &= ~skip
    if bad.any():
        idx = np.argwhere(bad)[:10]
        lines = [f"{name}{tuple(i)}: expected {expected[tuple(i)]!r} actual {actual[tuple(i)]!r}" for i in idx]
        raise AssertionError(f"{int(bad.sum())} elements of {name} exceed tolerance {tolerance}: max err "
                             f"{np.nanmax(err):.3e}\n" + "\n".join(lines))
    return float(np.nanmax(err)) if err.size else 0.0
