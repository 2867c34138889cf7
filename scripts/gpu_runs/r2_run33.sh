# r2 call 33: one-off fuzz of the tensor-core family (400 seeded random draws incl. transposes, D % 8 != 0, D > 128)
mkdir -p gpurun_out
timeout 2400 python scripts/fuzz_gpu.py --cases 400 --seed 11 --out gpurun_out/fuzz.jsonl 2>&1 | tail -30
