# r2 call 11: dK/dV at D = 128: exp2 fraction 0/4 vs 2/4 (the sweep's bf16 winner) under the reference policy (BF16 dO
# converted on chip) and bf16, interleaved, against the r1 build.
mkdir -p gpurun_out
for round in 0 1; do
for lib in r1 default; do
for table in builtin poly0; do
  if [ "$lib" = r1 ] && [ "$table" = poly0 ]; then continue; fi
  L=metal-flash-attention_b200/lib/libmfa_b200.so; [ "$lib" = r1 ] && L=metal-flash-attention_b200/lib/variants/libmfa_b200_r1.so
  T=""; [ "$table" = poly0 ] && T=$PWD/scripts/gpu_runs/r2_table_dkv_poly0.txt
  echo "== $round $lib $table"
  MFA_B200_LIBRARY=$PWD/$L MFA_B200_PARAMETER_FILE=$T python scripts/variant_sweep.py --child --configs 4096x128xREFx64,4096x128xBF16x64,4096x128xFP16x64 --kernels backwardKeyValue,backwardQuery --steps 20 | cut -c1-330
done; done; done > gpurun_out/dkv_d128_poly.txt 2>&1
cat gpurun_out/dkv_d128_poly.txt
