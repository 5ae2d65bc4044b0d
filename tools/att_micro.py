"""One non-local attention call in a loop (ncu / event timing).  usage: att_micro.py B Npos d dv [iters]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pretorched_x_b200 import ops

B, Npos, d, dv = [int(v) for v in sys.argv[1:5]]
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 20
dev = torch.device("cuda:0")
qkv = (torch.randn(B * Npos, 2 * d + dv, device=dev) * 0.3).half()
for _ in range(3):
    o = ops.nonlocal_attention(qkv, d, dv, B, Npos)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    o = ops.nonlocal_attention(qkv, d, dv, B, Npos)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("attention B=%d N=%d d=%d dv=%d: %.4f ms  %.1f TF/s" % (B, Npos, d, dv, ms, 2.0 * B * Npos * Npos * (d + dv) / ms / 1e9))
