"""One-off GPU fuzz of the tensor-core family against the CPU oracle: many seeded random (R, C, D, policy, transposes,
batch, lowPrecisionIntermediates) draws through forward + dQ + dK/dV, beyond the fixed lists of tests/test_tcgen05_stress.py.
Prints one line per failing case and a summary; exit status 1 if anything failed.  TEST INFRASTRUCTURE (imports oracle/).

Usage (GPU box):  python scripts/fuzz_gpu.py [--cases 300] [--seed 1] [--max-seq 900]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rel_rms(a, b):
    denom = float(np.sqrt(np.mean(b ** 2)))
    err = float(np.sqrt(np.mean((a - b) ** 2)))
    return err / denom if denom > 1e-12 else err


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-seq", type=int, default=900)
    ap.add_argument("--out", default=os.path.join("gpurun_out", "fuzz.jsonl"))
    args = ap.parse_args()
    import mfa_b200 as mfa
    import oracle
    from tests.attention_harness import run_attention
    KT, Op, P = mfa.AttentionKernelType, mfa.AttentionOperand, mfa.GEMMOperandPrecision
    rng = np.random.default_rng(args.seed)
    failures, backends = 0, {}
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as log:
        for case in range(args.cases):
            R = int(rng.integers(1, args.max_seq))
            C = int(rng.integers(1, args.max_seq))
            D = int(rng.integers(1, 257)) if rng.random() < 0.3 else int(rng.integers(1, 33)) * 8
            policy = ("bf16", "fp16", "reference")[int(rng.integers(0, 3))]
            lowMid = bool(rng.integers(0, 2))
            batch = int(rng.integers(1, 4))
            transpose = (False,) * 4
            if rng.random() < 0.4:
                mask = int(rng.integers(1, 16))
                transpose = tuple(bool(mask & (1 << i)) for i in range(4))
                if rng.random() < 0.8:   # mostly aligned (tensor cores); sometimes not (SIMT family)
                    R, C = max(8, R // 8 * 8), max(8, C // 8 * 8)
            desc = mfa.AttentionDescriptor()
            desc.lowPrecisionInputs = True
            desc.lowPrecisionIntermediates = lowMid
            if policy != "reference":
                desc.inputPrecisionOverride = P.BF16 if policy == "bf16" else P.FP16
            desc.matrixDimensions = (R, C, D)
            desc.transposeState = transpose
            desc.batchCount = batch
            tag = dict(case=case, R=R, C=C, D=D, policy=policy, lowMid=lowMid, batch=batch, transpose=list(transpose))
            try:
                kinds = tuple(desc.kernelDescriptor(t).backend.name for t in KT)
                backends[kinds] = backends.get(kinds, 0) + 1
                prec = desc.memoryPrecisions
                nets = [oracle.Network(R, C, D, seed=7919 * case + b, threads=8).round_inputs(int(prec[Op.Q]), int(prec[Op.dO]))
                        for b in range(batch)]
                inputs = {getattr(Op, k): np.stack([getattr(n, k) for n in nets]) if batch > 1 else getattr(nets[0], k)
                          for k in ("Q", "K", "V", "dO")}
                out = run_attention(desc, None, inputs=inputs)
                bf16 = policy == "bf16"
                worst = {}
                for b, n in enumerate(nets):
                    pick = (lambda a: a[b]) if batch > 1 else (lambda a: a)
                    O, L = n.inferenceAttention(with_L=True)
                    ref = {"O": O, "dV": n.derivativeV(), "dK": n.derivativeK(), "dQ": n.derivativeQ()}
                    for name, expected in ref.items():
                        worst[name] = max(worst.get(name, 0.0), rel_rms(pick(out[name]), expected))
                    worst["L"] = max(worst.get("L", 0.0), float(np.abs(pick(out["L"]) - L).max()))
                bound = 4e-3 if bf16 else 1.5e-3
                if lowMid:
                    bound = max(bound, 6e-3)
                if min(R, C, D) < 16:
                    bound *= 1.5
                ok = all(worst[k] <= bound for k in ("O", "dV", "dK", "dQ")) and worst["L"] <= (7e-3 if lowMid else 1e-3)
                tag.update(backends=list(kinds), worst=worst, ok=ok)
            except Exception as exc:  # a crash is a failure too
                tag.update(error=repr(exc)[:300], ok=False)
            if not tag["ok"]:
                failures += 1
                print("FAIL", json.dumps(tag), flush=True)
            log.write(json.dumps(tag) + "\n")
    print(f"{args.cases} cases, {failures} failures; backend mix (forward, dQ, dK/dV): "
          + ", ".join(f"{'/'.join(k)} x{v}" for k, v in sorted(backends.items(), key=lambda kv: -kv[1])), flush=True)
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
