mkdir -p gpurun_out
timeout 120 python scripts/pcie_probe.py > gpurun_out/pcie_probe.log 2>&1
cat gpurun_out/pcie_probe.log | cut -c1-400
