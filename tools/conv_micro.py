"""One convolution layer in a loop (for ncu / event timing of a single kernel shape), optionally checked against the CUDA-core
cross-check kernel.
usage: conv_micro.py N Cin T H W K kt kh kw st sh sw [iters] [--check] [--residual]"""
import os, sys
import torch, torch.nn as nn
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pretorched_x_b200 import ops, engine

flags = [v for v in sys.argv[1:] if v.startswith("--")]
nums = [v for v in sys.argv[1:] if not v.startswith("--")]
a = [int(v) for v in nums[:12]]
iters = int(nums[12]) if len(nums) > 12 else 20
N, Cin, T, H, W, K, kt, kh, kw, st, sh, sw = a
dev = torch.device("cuda:0")
torch.manual_seed(0)
conv = nn.Conv3d(Cin, K, (kt, kh, kw), stride=(st, sh, sw), padding=(kt // 2, kh // 2, kw // 2), bias=False).to(dev)
bn = nn.BatchNorm3d(K).eval().to(dev)
with torch.no_grad():
    bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5); bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1)
x = ops.from_ncdhw(torch.randn(N, Cin, T, H, W, device=dev))
res = None
with torch.no_grad():
    y = engine.conv_bn_act(conv, bn, x, relu=True)
    if "--residual" in flags:
        res = ops.Act(torch.randn_like(y.data.float()).half(), y.N, y.T, y.H, y.W, y.C)
        res.data[:, y.C:] = 0
    for _ in range(3):
        y = engine.conv_bn_act(conv, bn, x, residual=res, relu=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        y = engine.conv_bn_act(conv, bn, x, residual=res, relu=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    M = y.data.shape[0]
    fl = 2.0 * M * K * Cin * kt * kh * kw
    msg = "conv %s%s: %.4f ms  %.1f TF/s  M=%d" % (a, " +res" if res is not None else "", ms, fl / ms / 1e9, M)
    if "--check" in flags:
        want = engine.conv_bn_act(conv, bn, x, residual=res, relu=True, simt=True)
        err = (y.data.float() - want.data.float()).abs().max().item() / max(want.data.float().abs().max().item(), 1e-6)
        msg += "  rel err vs CUDA-core kernel %.2e %s" % (err, "OK" if err < 2e-3 else "MISMATCH")
    print(msg, flush=True)
