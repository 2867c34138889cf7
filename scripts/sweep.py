"""Regenerates the tensor-core family's parameter tables from measurements on the current GPU -- the analogue of the
reference's parameter sweeps (Tests/FlashAttentionTests/Attention/SquareAttentionTest.swift:29-132, which produced the
tables in AttentionDescriptor+Parameters.swift:106-285).

The tables are data: every candidate is a parameter FILE (MFA_B200_PARAMETER_FILE) read by an otherwise identical
child process, so what is timed is exactly what a user of that table would launch.  Swept per (kernel type, head
dimension bucket):
  * the exp2-on-the-FMA-pipe fraction (0 .. mfa_max_exp2_fma_quarters quarters of the element pairs) at a large batch;
  * the small-grid split policy (minimum blocks per range x maximum ranges) for a single head.
Writes the winning tables to --out (default metal-flash-attention_b200/parameters/b200.txt, loadable through
MFA_B200_PARAMETER_FILE and the source of the built-in defaults in csrc/descriptor.cpp) and one JSON line per timed
candidate to --log.

Usage (GPU box):  python scripts/sweep.py [--quick]
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SECTIONS = {"forward": "[forward]", "backwardQuery": "[backwardQuery]", "backwardKeyValue": "[backwardKeyValue]"}
RESIDENT = {"forward": "Q, O", "backwardQuery": "Q, dO, dQ", "backwardKeyValue": "K, V, dV, dK"}
PAR = {"forward": 256, "backwardQuery": 128, "backwardKeyValue": 128}


def table_text(kind, rows):
    """rows: {bucket: (quarters, min_blocks, max_splits)} for the buckets 64 and 128 (+ the fixed 256 forward row)."""
    lines = []
    for bucket in (64, 128):
        q, mb, ms = rows[bucket]
        lines.append(f"| {bucket:<3d} | {PAR[kind]} | 128 | {bucket:<3d} | {RESIDENT[kind]} | {q} | {mb} | {ms} |")
    # rows that are not swept: one compiled configuration each (the D <= 256 forward; the wide-head backward kernels,
    # whose split policy follows the 128 row's)
    if kind == "forward":
        lines.append("| 256 | 128 | 128 | 256 | Q, O | 0 | 0 | 1 |")
    else:
        _, mb, ms = rows[128]
        lines.append(f"| 256 | 128 | 64  | 256 | {RESIDENT[kind]} | 0 | {mb} | {ms} |")
    return "\n".join(lines) + "\n"


def transposed_text(kind, rows):
    """The layout-generic kernels' tables (transposed operands)."""
    if kind == "forward":
        return "| 128 | 128 | 128 | 128 | Q, O | 0 | 0 | 1 |\n| 256 | 128 | 128 | 256 | Q, O | 0 | 0 | 1 |\n"
    _, mb, ms = rows[128]
    return "".join(f"| {b:<3d} | 128 | 64  | {b:<3d} | {RESIDENT[kind]} | 0 | {mb} | {ms} |\n" for b in (64, 128, 256))


def write_file(path, tables):
    with open(path, "w") as f:
        f.write("# B200 parameter tables of the tcgen05 kernel family (scripts/sweep.py).  Columns: max head dimension |\n"
                "# parallelization | traversal | head block | resident operands | exp2 on the FMA pipe (quarters of the\n"
                "# element pairs) | minimum blocks per split range (0 = never split) | maximum split ranges\n")
        for kind, rows in tables.items():
            f.write(SECTIONS[kind] + "\n" + table_text(kind, rows))
            f.write(SECTIONS[kind][:-1] + ".transposed]\n" + transposed_text(kind, rows))


def child(spec):
    sys.path.insert(0, ROOT)
    import torch
    import mfa_b200 as mfa
    from scripts.bench_configs import run
    P = mfa.GEMMOperandPrecision
    out = {}
    for N, D, H in spec:
        r = run(N, D, P.BF16, H, steps=20 if H > 1 else 200)
        out[f"{N}x{D}x{H}"] = {k: {"tflops": v["tflops"], "ms": v["ms"], "kernel": v["kernel"]}
                               for k, v in r.items() if isinstance(v, dict)}
    print(json.dumps(out), flush=True)


def run_candidate(tables, spec, log, tag):
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        path = f.name
    write_file(path, tables)
    env = dict(os.environ, MFA_B200_PARAMETER_FILE=path)
    p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", json.dumps(spec)], env=env,
                       capture_output=True, text=True)
    os.unlink(path)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    if not lines:
        raise RuntimeError(p.stderr[-2000:])
    result = json.loads(lines[-1])
    log.write(json.dumps({"candidate": tag, "tables": {k: {str(b): list(v) for b, v in rows.items()} for k, rows in tables.items()},
                          "result": result}) + "\n")
    log.flush()
    return result


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "metal-flash-attention_b200", "parameters", "b200.txt"))
    ap.add_argument("--log", default=os.path.join(ROOT, "gpurun_out", "parameter_sweep.jsonl"))
    ap.add_argument("--quick", action="store_true", help="fewer split-policy candidates")
    args = ap.parse_args()
    if args.child:
        child(json.loads(args.child))
        return
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    os.makedirs(os.path.dirname(args.log), exist_ok=True)
    kinds = ("forward", "backwardQuery", "backwardKeyValue")
    max_q = {"forward": 2, "backwardQuery": 3, "backwardKeyValue": 3}
    default_split = {"forward": (4, 16), "backwardQuery": (2, 8), "backwardKeyValue": (2, 8)}
    best = {k: {64: (0,) + default_split[k], 128: (0,) + default_split[k]} for k in kinds}
    with open(args.log, "w") as log:
        # ---- exp2 fraction, throughput regime: 64 heads (D = 128) / 128 heads (D = 64) ----
        big = [(4096, 128, 64), (2048, 64, 128)]
        score = {k: {64: {}, 128: {}} for k in kinds}
        for q in range(0, 4):
            tables = {k: {b: (min(q, max_q[k]),) + default_split[k] for b in (64, 128)} for k in kinds}
            res = run_candidate(tables, big, log, f"exp2={q}/4")
            for k in kinds:
                if q <= max_q[k]:
                    score[k][128][q] = res["4096x128x64"][k]["tflops"]
                    score[k][64][q] = res["2048x64x128"][k]["tflops"]
        for k in kinds:
            for b in (64, 128):
                q = max(score[k][b], key=score[k][b].get)
                best[k][b] = (q,) + default_split[k]
                print(f"{k:17s} D<={b:3d}: exp2 on FMA pipe {q}/4   (TFLOP/s by fraction: {score[k][b]})", flush=True)
        # ---- split policy, latency regime: one head ----
        small = [(4096, 128, 1), (2048, 64, 1)]
        candidates = [(2, 8), (4, 8), (4, 16), (8, 4)] if args.quick else [(2, 4), (2, 8), (4, 4), (4, 8), (4, 16), (8, 4), (8, 8), (0, 1)]
        lat = {k: {64: {}, 128: {}} for k in kinds}
        for mb, ms in candidates:
            tables = {k: {b: (best[k][b][0], mb, ms) for b in (64, 128)} for k in kinds}
            res = run_candidate(tables, small, log, f"split=({mb},{ms})")
            for k in kinds:
                lat[k][128][(mb, ms)] = res["4096x128x1"][k]["ms"]
                lat[k][64][(mb, ms)] = res["2048x64x1"][k]["ms"]
        for k in kinds:
            for b in (64, 128):
                mb, ms = min(lat[k][b], key=lat[k][b].get)
                best[k][b] = (best[k][b][0], mb, ms)
                print(f"{k:17s} D<={b:3d}: split policy min blocks {mb}, max ranges {ms}   (us by policy: "
                      f"{ {str(c): round(v * 1e3, 2) for c, v in lat[k][b].items()} })", flush=True)
    write_file(args.out, best)
    print(open(args.out).read())


if __name__ == "__main__":
    main()
