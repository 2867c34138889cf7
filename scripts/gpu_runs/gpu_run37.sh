mkdir -p gpurun_out
cat > /tmp/prof_bwd.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from scripts.bench_configs import run
print(run(2048, 64, None, 128, steps=1))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_backward -s 3 -c 2 -f -o gpurun_out/r1_bwd_d64 python /tmp/prof_bwd.py > gpurun_out/ncu_bwd_d64.log 2>&1
ls -la gpurun_out/r1_bwd_d64.ncu-rep | cut -c30-
