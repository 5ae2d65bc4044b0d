"""Throughput of the other BASELINE.json configs (parity-test cases, not bench lines): r2plus1d34 B=16 32x112x112 and
nonlocalresnet3d50 B=8 32x224x224 per GPU, CUDA-graph replay, device-resident input.  Prints one JSON line each."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pretorched_x_b200 as P
from pretorched_x_b200 import ops
from pretorched_x_b200.graph import GraphedForward
from oracle import functional as OF

dev = torch.device("cuda:0")
CASES = [("r2plus1d34", dict(num_classes=400), (16, 3, 32, 112, 112), 51.48),
         ("nonlocalresnet3d50", dict(pretrained=None), (8, 3, 32, 224, 224), 262.22),
         ("resnet18", dict(num_classes=1000, pretrained=None), (256, 3, 224, 224), 3.63),
         ("slowfast.resnet50", dict(mode="sf", num_classes=400), (8, 3, 64, 224, 224), 0.0)]
if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
    CASES = [c for c in CASES if c[0] in sys.argv[1].split(",")]
for arch, kw, shape, gflop in CASES:
    torch.manual_seed(0)
    f = P
    for part in arch.split("."):
        f = getattr(f, part)
    m = f(**kw)
    OF.randomize_bn_(m, 1)
    m = m.eval().to(dev)
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1)).to(dev)
    with torch.no_grad():
        m(x)
        with ops.profile() as prof:
            m(x)
    g = GraphedForward(m, x)
    for _ in range(3):
        g()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 20
    e0.record()
    for _ in range(steps):
        g()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    rate = shape[0] / (ms / 1e3)
    agg = {}
    lay = {}
    for r in prof.rows:
        a = agg.setdefault(r["kind"], [0.0, 0]); a[0] += r["ms"]; a[1] += 1
        b = lay.setdefault(r["desc"], [0.0, 0, 0.0]); b[0] += r["ms"]; b[1] += 1; b[2] += r["flops"]
    if "--layers" in sys.argv:
        for d, v in sorted(lay.items(), key=lambda kv: -kv[1][0])[:14]:
            print("   %-52s n=%2d %7.3f ms %7.1f TF/s" % (d, v[1], v[0], v[2] / v[0] / 1e9 if v[0] else 0), file=sys.stderr)
    print(json.dumps({"arch": arch, "input": shape, "ms_per_step": ms, "samples_per_s": rate,
                      "tflops_algorithmic": gflop * rate / 1e3,
                      "eager_ms_by_kind": {k: round(v[0], 3) for k, v in agg.items()}, "launches": len(prof.rows)}), flush=True)
    del g, m, x
    torch.cuda.empty_cache()
