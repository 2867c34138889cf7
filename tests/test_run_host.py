"""mfa_attention_run_host (the e2e entry point bench.py times): host buffers in, host buffers out.  Checked against
the CPU oracle and against the device-pointer path, for a single problem and for batches that the library cuts into
chunks rotating over three streams (uploads, kernels and downloads of neighbouring chunks overlap)."""
import numpy as np
import pytest


def _host_run(desc, nets, types, pinned):
    import torch
    import mfa_b200 as mfa
    import oracle
    Op = mfa.AttentionOperand
    prec = desc.memoryPrecisions
    R, C, D = desc.matrixDimensions
    B = len(nets)
    host = {}
    for op, name in ((Op.Q, "Q"), (Op.K, "K"), (Op.V, "V"), (Op.dO, "dO")):
        raw = oracle.encode(np.stack([getattr(n, name) for n in nets]).astype(np.float32), int(prec[op]))
        t = torch.from_numpy(raw.view(np.int16) if raw.dtype == np.uint16 else raw)
        host[op] = t.pin_memory() if pinned else t.clone()
    shapes = {Op.O: (B, R, D), Op.L: (B, R), Op.D: (B, R), Op.dQ: (B, R, D), Op.dK: (B, C, D), Op.dV: (B, C, D)}
    for op, shape in shapes.items():
        dt = torch.float32 if prec[op] == mfa.GEMMOperandPrecision.FP32 else torch.int16
        t = torch.full(shape, float("nan") if dt == torch.float32 else -1, dtype=dt)
        host[op] = t.pin_memory() if pinned else t
    desc.runHost(types, {op: t.data_ptr() for op, t in host.items()}, device=torch.cuda.current_device())
    out = {}
    for op in shapes:
        a = host[op].numpy()
        out[op.name] = a if a.dtype == np.float32 else oracle.decode(a.view(np.uint16), int(prec[op]))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("batch,R,C,D,bf16,pinned", [(1, 200, 333, 64, True, True), (5, 256, 256, 128, True, True),
                                                     (19, 130, 70, 64, False, True), (3, 64, 48, 24, None, False),
                                                     (40, 1024, 1024, 128, True, True)])
def test_run_host_matches_oracle_and_device_path(batch, R, C, D, bf16, pinned):
    import mfa_b200 as mfa
    import oracle
    from tests.attention_harness import run_attention, LOG2E

    KT = mfa.AttentionKernelType
    desc = mfa.AttentionDescriptor()
    desc.lowPrecisionInputs = bf16 is not None          # None: the FP32 family
    if bf16 is not None:
        desc.inputPrecisionOverride = mfa.GEMMOperandPrecision.BF16 if bf16 else mfa.GEMMOperandPrecision.FP16
    desc.matrixDimensions = (R, C, D)
    desc.transposeState = (False, False, False, False)
    desc.batchCount = batch
    prec = desc.memoryPrecisions
    nets = [oracle.Network(R, C, D, seed=100 + b, threads=8).round_inputs(int(prec[mfa.AttentionOperand.Q]),
                                                                             int(prec[mfa.AttentionOperand.dO]))
            for b in range(batch)]
    out = _host_run(desc, nets, list(KT), pinned)

    # device-pointer path on the same inputs: the chunked host path must reproduce it bit for bit (chunking only
    # changes which launch a head belongs to, never its arithmetic) -- except where split-KV engages for small grids
    inputs = {getattr(mfa.AttentionOperand, k): np.stack([getattr(n, k) for n in nets]) for k in ("Q", "K", "V", "dO")}
    dev = run_attention(desc, None, inputs=inputs, return_raw=True)
    for name in ("O", "L", "D", "dQ", "dK", "dV"):
        a, b = out[name].reshape(dev[name].shape), dev[name]
        assert np.isfinite(a).all(), name
        np.testing.assert_allclose(a, b, rtol=2e-3 if bf16 is not None else 1e-5, atol=2e-3 if bf16 is not None else 1e-5,
                                   err_msg=name)

    # and the oracle, on a few of the heads
    tol = 2e-5 if bf16 is None else 5e-2
    for b in sorted({0, batch // 2, batch - 1}):
        n = nets[b]
        O, L = n.inferenceAttention(with_L=True)
        assert np.abs(out["O"][b] - O).max() <= tol
        assert np.abs(out["L"][b] / np.float32(LOG2E) - L).max() <= (2e-5 if bf16 is None else 7e-3)
        for name, expected in (("dV", n.derivativeV()), ("dK", n.derivativeK()), ("dQ", n.derivativeQ())):
            assert np.abs(out[name][b] - expected).max() <= tol, (name, b)


@pytest.mark.gpu
def test_run_host_forward_only_leaves_other_outputs_untouched():
    import mfa_b200 as mfa
    import oracle
    desc = mfa.AttentionDescriptor()
    desc.lowPrecisionInputs = True
    desc.inputPrecisionOverride = mfa.GEMMOperandPrecision.BF16
    desc.matrixDimensions = (300, 300, 128)
    desc.transposeState = (False, False, False, False)
    desc.batchCount = 12
    nets = [oracle.Network(300, 300, 128, seed=b, threads=8).round_inputs(oracle.BF16) for b in range(12)]
    out = _host_run(desc, nets, [mfa.AttentionKernelType.forward], True)
    assert np.isfinite(out["O"]).all() and np.isnan(out["dQ"]).all() and np.isnan(out["dK"]).all()
    O = nets[7].inferenceAttention()
    assert np.abs(out["O"][7] - O).max() <= 5e-3
