#!/bin/bash
# Depth-first schedule on the GPU box: parity tests of the chunked walks, then the CUDA-graph-timed sweeps (tools/dfs_sweep.py).
O=gpurun_out/dfs; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; tail -15 $O/pytest_gpu.txt
for w in resnet3d50 r2plus1d34 nonlocal50 resnet18 biggan256; do
  timeout 300 python tools/dfs_sweep.py $w > $O/sweep_$w.txt 2>&1
done
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
tail -n 30 $O/sweep_*.txt
