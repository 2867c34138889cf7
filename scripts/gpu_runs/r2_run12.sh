# r2 call 12 (4 GPUs, all on one socket in this pool's boxes): the bench line under torchrun -- e2e scaling with
# NUMA-local host buffers, configs[4] with batched NCCL scatter / gather.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_4gpu.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench_n4.log 2> gpurun_out/bench_n4.err; echo "exit $?" >> gpurun_out/bench_n4.log
grep "^{" gpurun_out/bench_n4.log | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'])
print('sustained',d['sustained']['tflops_per_gpu'],d['sustained']['clocks'])
c=d['config5']; print({k:c[k] for k in ('kernel_ms','scatter_ms','gather_ms','scatter_gbs','gather_gbs','with_scatter_gather_ms')})
"
tail -n 3 gpurun_out/bench_n4.err | cut -c1-300
