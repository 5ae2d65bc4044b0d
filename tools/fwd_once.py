"""Two eager forwards of one bench workload (for `ncu --metrics gpu__time_duration.sum`: the second forward's launches are the
per-kernel device times).  usage: fwd_once.py workload [batch]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pretorched_x_b200 import _lib
spec = bench.WORKLOADS[sys.argv[1]]
B = int(sys.argv[2]) if len(sys.argv) > 2 else spec["batch"]
dev = torch.device("cuda:0")
m = bench.build_ours(spec).to(dev)
x = torch.randn((B,) + spec["sample"], generator=torch.Generator().manual_seed(1)).to(dev)
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    c0 = _lib.launch_count()
    m(x); torch.cuda.synchronize()
print("launches per forward:", _lib.launch_count() - c0)
