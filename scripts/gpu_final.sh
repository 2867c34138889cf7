# Round-end validation: every GPU test, smoke, headline bench, the other configs, ncu launch list and full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "exit $?" >> gpurun_out/smoke.log
tail -n 3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2>&1; echo "exit $?" >> gpurun_out/bench.log
grep "^{" gpurun_out/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['tflops'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['single_head'], d['gpu_launches'], d['clocks'], d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1
tail -n 1 gpurun_out/bench_reference.log | cut -c1-250
timeout 300 python scripts/bench_configs.py > gpurun_out/bench_configs.log 2>&1; echo "exit $?" >> gpurun_out/bench_configs.log
python - <<'PY'
import json
for l in open("gpurun_out/bench_configs.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["N"], d["D"], d["dtype"][:12], d["heads"], {k: (v["ms"], v["tflops"]) for k, v in d.items() if isinstance(v, dict)})
    else:
        print(l.strip()[:300])
PY
timeout 300 python scripts/bench_single.py > gpurun_out/bench_single.log 2>&1; echo "exit $?" >> gpurun_out/bench_single.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_forward_tcgen05 -s 4 -c 1 -f -o gpurun_out/r1_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_fwd.log 2>&1
cat > /tmp/prof_d256.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from scripts.bench_configs import run
import mfa_b200 as mfa
print(run(8192, 256, mfa.GEMMOperandPrecision.BF16, 16, steps=1))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_forward_d256 -s 3 -c 1 -f -o gpurun_out/r1_fwd_d256 python /tmp/prof_d256.py > gpurun_out/ncu_fwd_d256.log 2>&1
cat > /tmp/prof_bwd.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from scripts.bench_configs import run
import mfa_b200 as mfa
N, D, prec, H = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
p = {"bf16": mfa.GEMMOperandPrecision.BF16, "fp16": mfa.GEMMOperandPrecision.FP16, "ref": None}[prec]
print(run(N, D, p, H, steps=1))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_backward -s 3 -c 2 -f -o gpurun_out/r1_bwd_d128 python /tmp/prof_bwd.py 4096 128 bf16 32 > gpurun_out/ncu_bwd_d128.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_backward -s 3 -c 2 -f -o gpurun_out/r1_bwd_d64 python /tmp/prof_bwd.py 2048 64 ref 128 > gpurun_out/ncu_bwd_d64.log 2>&1
ls -la gpurun_out/*.ncu-rep | cut -c30-
