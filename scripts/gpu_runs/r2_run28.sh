# r2 call 28 (8 GPUs): BASELINE configs[4] as written -- 2048 problems, 256 per GPU, NCCL scatter / kernels / gather --
# NUMA-local host buffers, configs[4] with batched NCCL scatter / gather.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8gpu.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "exit $?" >> gpurun_out/bench_n8.log
grep "^{" gpurun_out/bench_n8.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
print('sustained',d['sustained']['tflops_per_gpu'],d['sustained']['clocks'])
c=d['config5']; print({k:c[k] for k in ('kernel_ms','scatter_ms','gather_ms','scatter_gbs','gather_gbs','with_scatter_gather_ms')})
"
tail -n 3 gpurun_out/bench_n8.err | cut -c1-300
