"""One 2-D convolution of the BigGAN path in a loop (for ncu / event timing of a single kernel shape).
usage: gan_micro.py N C H W K k [up=0|1] [sample_affine=0|1] [iters] [in_affine=0|1]      (k = 1 or 3; H, W = INPUT size)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pretorched_x_b200 import ops
from pretorched_x_b200.ops import Act

a = [int(v) for v in sys.argv[1:7]]
N, C, H, W, K, k = a
up = int(sys.argv[7]) if len(sys.argv) > 7 else 0
saff = int(sys.argv[8]) if len(sys.argv) > 8 else 0
iters = int(sys.argv[9]) if len(sys.argv) > 9 else 20
inaff = int(sys.argv[10]) if len(sys.argv) > 10 else 0
dev = torch.device("cuda:0")
w = torch.randn(K, C, k, k, device=dev) / (k * C ** 0.5)
pc = ops.PackedConv(w, torch.randn(K, device=dev), None, (1, 1, 1), (0, k // 2, k // 2), in_pitch=C, upsample=bool(up))
x = Act(torch.randn(N * H * W, C, device=dev).half(), N, 1, H, W, C)
aff = torch.randn(N, 2 * K + 64, device=dev)
sa = (aff[:, :K], aff[:, K + 64:]) if saff else None
aff_in = torch.randn(N, 2 * C + 64, device=dev)
ia = (aff_in[:, :C], aff_in[:, C + 64:]) if inaff else None      # bn1 + ReLU on the A operand (1x1 convolutions)
for _ in range(3):
    y = ops.conv(x, pc, relu=True, sample_affine=sa, in_affine=ia)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    y = ops.conv(x, pc, relu=True, sample_affine=sa, in_affine=ia)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
M = y.data.shape[0]
fl = 2.0 * M * K * C * k * k
print("conv2d %s up=%d saff=%d: %.4f ms  %.1f TF/s algorithmic  M=%d  out %.0f GB/s" % (a, up, saff, ms, fl / ms / 1e9, M, M * K * 2 / ms / 1e6))
