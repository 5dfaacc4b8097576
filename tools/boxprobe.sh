#!/bin/bash
# What the GPU box offers the host side of the tile loops: cores, memory, tmpfs, topology.
echo "== nproc: $(nproc)"; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" 
free -g | head -2
df -h /dev/shm /tmp | cat
nvidia-smi --query-gpu=index,name,pcie.link.gen.current,pcie.link.width.current --format=csv
nvidia-smi topo -m 2>/dev/null | head -20
ulimit -l
