#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2g4; mkdir -p $O
nvidia-smi topo -m 2>&1 | head -12
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29651 tools/pcie_probe.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee $O/pcie_probe_n4.txt
