"""One convolution layer in a loop (for ncu / event timing of a single kernel shape).
usage: conv_micro.py N Cin T H W K kt kh kw st sh sw [iters]"""
import os, sys
import torch, torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pretorched_x_b200 import ops, engine

a = [int(v) for v in sys.argv[1:13]]
iters = int(sys.argv[13]) if len(sys.argv) > 13 else 20
N, Cin, T, H, W, K, kt, kh, kw, st, sh, sw = a
dev = torch.device("cuda:0")
conv = nn.Conv3d(Cin, K, (kt, kh, kw), stride=(st, sh, sw), padding=(kt // 2, kh // 2, kw // 2), bias=False).to(dev)
bn = nn.BatchNorm3d(K).eval().to(dev)
x = ops.from_ncdhw(torch.randn(N, Cin, T, H, W, device=dev))
for _ in range(3):
    y = engine.conv_bn_act(conv, bn, x, relu=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    y = engine.conv_bn_act(conv, bn, x, relu=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
M = y.data.shape[0]
fl = 2.0 * M * K * Cin * kt * kh * kw
print("conv %s: %.4f ms  %.1f TF/s  M=%d" % (a, ms, fl / ms / 1e9, M))
