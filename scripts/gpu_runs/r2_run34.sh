# r2 call 34: second fuzz run, longer sequences (split paths, many blocks per item), another seed
mkdir -p gpurun_out
timeout 2400 python scripts/fuzz_gpu.py --cases 500 --seed 23 --max-seq 2200 --out gpurun_out/fuzz2.jsonl 2>&1 | tail -30
