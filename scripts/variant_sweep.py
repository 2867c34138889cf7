"""A/B timing of library variants on one box: every variant (a libmfa_b200 build, see csrc/Makefile VARIANT=...) runs the
same configs in its own process (MFA_B200_LIBRARY selects the .so), the whole list is walked `--rounds` times so that
clock / power drift shows up as spread instead of as a winner.  One JSON line per (round, variant).

Usage (GPU box):  python scripts/variant_sweep.py --variants default,r1,poly1 --configs 4096x128xBF16x64,2048x64xFP16x128
                  [--kernels forward] [--rounds 2]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(configs, kernels, steps):
    sys.path.insert(0, ROOT)
    import torch  # noqa: F401
    import mfa_b200 as mfa
    from scripts.bench_configs import run
    P = mfa.GEMMOperandPrecision
    res = {"lib": os.path.basename(mfa.library_path())}
    for spec in configs:
        n, d, prec, h = spec.split("x")
        precision = None if prec == "REF" else P[prec]
        r = run(int(n), int(d), precision, int(h), steps=steps)
        res[spec] = {k: v["tflops"] for k, v in r.items() if isinstance(v, dict) and (not kernels or k in kernels)}
        res.setdefault("clocks", {})[spec] = {k: (v["clocks"]["sm_mhz"], v["clocks"]["reasons"])
                                              for k, v in r.items() if isinstance(v, dict)}
    print(json.dumps(res), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="default")
    ap.add_argument("--configs", default="4096x128xBF16x64")
    ap.add_argument("--kernels", default="")
    ap.add_argument("--rounds", type=int, default=2)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    configs = args.configs.split(",")
    kernels = [k for k in args.kernels.split(",") if k]
    if args.child:
        child(configs, kernels, args.steps)
        return
    lib_dir = os.path.join(ROOT, "metal-flash-attention_b200", "lib")
    for rnd in range(args.rounds):
        for v in args.variants.split(","):
            path = os.path.join(lib_dir, "libmfa_b200.so") if v == "default" else \
                os.path.join(lib_dir, "variants", f"libmfa_b200_{v}.so")
            env = dict(os.environ, MFA_B200_LIBRARY=path)
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--configs", args.configs,
                                "--kernels", args.kernels, "--steps", str(args.steps)], env=env, capture_output=True,
                               text=True)
            line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            if line:
                d = json.loads(line[-1])
                d["variant"], d["round"] = v, rnd
                print(json.dumps(d), flush=True)
            else:
                print(json.dumps({"variant": v, "round": rnd, "error": (p.stderr or p.stdout)[-400:]}), flush=True)


if __name__ == "__main__":
    main()
