# r2 call 8 (2 GPUs): the bench line under torchrun (NCCL scatter/gather of configs[4], NUMA-local e2e), the single-process
# two-GPU host test, topology.
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_2gpu.txt 2>&1
timeout 600 python -m pytest tests/test_run_host.py -m gpu -q -x -k "two_gpus or numa" > gpurun_out/pytest_2gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_2gpu.log
tail -n 5 gpurun_out/pytest_2gpu.log
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "exit $?" >> gpurun_out/bench_n2.log
grep "^{" gpurun_out/bench_n2.log | cut -c1-4500
grep -i "NCCL INFO.*\(P2P\|via\|nranks\|comm 0x\)" gpurun_out/bench_n2.err gpurun_out/bench_n2.log | head -12 | cut -c1-250
tail -n 5 gpurun_out/bench_n2.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1
grep "^{" gpurun_out/bench_ref_n2.log | cut -c1-300
