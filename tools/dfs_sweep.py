#!/usr/bin/env python
"""CUDA-graph-timed sweep of the trunk schedule (engine.run_trunk: breadth-first vs depth-first in L2-sized clip chunks).

    python tools/dfs_sweep.py resnet3d50 [spec ...]      # spec: off | units:clips[,units:clips...]  (biggan256: res:images,...)

Prints one line per spec: ms per forward (min and median of 3 rounds of `--reps` graph replays), clips/s, and the maximum
deviation of the logits from the breadth-first walk (0 = bit-identical).  Same model / input construction as bench.py.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from pretorched_x_b200 import engine  # noqa: E402
from pretorched_x_b200.graph import GraphedForward  # noqa: E402

DEFAULT_SPECS = {
    "resnet3d50": ["off", "1:1", "1:2", "1:4", "4:1", "4:2", "4:4", "4:8", "5:1", "5:2", "5:4", "5:2,3:8", "5:2,4:8",
                   "5:1,4:8", "5:4,4:8", "5:2,4:16", "5:2,4:4", "8:4", "8:8", "5:2,4:8,5:16", "4:2,5:8", "off"],
    "nonlocal50": ["off", "4:1", "4:2", "5:1", "5:2", "5:1,4:2", "5:1,4:4", "9:1", "9:2", "off"],
    "resnet18": ["off", "1:16", "1:32", "3:8", "3:16", "3:32", "3:64", "4:16", "4:32", "5:32", "5:64", "4:16,1:64", "4:32,5:64", "off"],
    "r2plus1d34": ["off", "1:1", "1:2", "1:4", "4:2", "4:4", "1:1,3:4", "1:2,3:4", "1:2,3:8", "off"],
    "trn": ["off", "5:8", "5:16", "5:8,4:32", "5:16,4:64", "off"],
    # generator: res:images levels (biggan_engine.dfs_plan)
    "biggan256": ["off", "256:2", "256:4", "256:8", "256:16", "128:4", "128:8", "128:16", "128:8,256:4", "128:16,256:8",
                  "64:16,128:8,256:4", "64:32,128:16,256:8", "64:32,128:8,256:4", "32:64,64:32,128:8,256:4", "64:8", "64:16", "off"],
}


def apply_spec(spec):
    engine.set_dfs(spec)


def time_graph(gf, reps):
    times = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            gf.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) / reps)
    return sorted(times)


def sweep_biggan(args, dev):
    import pretorched_x_b200 as P
    from pretorched_x_b200 import biggan_engine
    from oracle import biggan as OB
    B = args.batch or bench.BIGGAN_BATCH
    model, _, _, _ = OB.build_case(P.biggan_deep, bench.BIGGAN_RES, bench.BIGGAN_CH, bench.BIGGAN_CLASSES, 4, init="ortho")
    model = model.to(dev)
    z, lab = OB.seeded_inputs(B, bench.BIGGAN_CLASSES, 1000)
    z, lab = z.to(dev), lab.to(dev)
    mods = [(i, blk) for i, stage in enumerate(model.blocks) for blk in stage]
    ref = None
    for spec in (args.specs or DEFAULT_SPECS["biggan256"]):
        biggan_engine.set_dfs(spec)
        try:
            gf = GraphedForward(model, (z, lab), warmup=1, out_dtype=torch.float16)
            out = gf().float().clone()
            if ref is None:
                ref = out
            times = time_graph(gf, max(3, args.reps // 3))
            print("%-12s %-28s plan=%-40s %8.3f ms (median %8.3f)  %9.0f /s  max dev %.2e rms %.2e" % (
                "biggan256", spec, biggan_engine.dfs_plan(model, B, mods), times[0], times[1], B / times[0] * 1e3,
                (out - ref).abs().max().item(), (out - ref).pow(2).mean().sqrt().item()), flush=True)
            del gf, out
        except Exception as e:      # noqa: BLE001
            print("%-12s %-28s FAILED: %r" % ("biggan256", spec, e), flush=True)
        torch.cuda.empty_cache()
    biggan_engine.set_dfs("off")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("specs", nargs="*")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    if args.workload == "biggan256":
        return sweep_biggan(args, dev)
    spec_w = bench.WORKLOADS[args.workload]
    model = bench.build_ours(spec_w).to(dev)
    B = args.batch or spec_w["batch"]
    g = torch.Generator().manual_seed(1000)
    x = torch.randn((B,) + spec_w["sample"], generator=g).to(dev)
    specs = args.specs or DEFAULT_SPECS[args.workload]
    ref = None
    for spec in specs:
        apply_spec(spec)
        plan = engine.dfs_plan()
        try:
            gf = GraphedForward(model, x, warmup=1)
            out = gf().float().clone()
            if ref is None:
                ref = out
            dev_max = (out - ref).abs().max().item() / ref.abs().max().item()
            times = time_graph(gf, args.reps)
            print("%-12s %-28s plan=%-28s %8.3f ms (median %8.3f)  %9.0f /s  dev %.2e" % (
                args.workload, spec, plan, times[0], times[1], B / times[0] * 1e3, dev_max), flush=True)
            del gf
        except Exception as e:      # noqa: BLE001 - a sweep keeps going
            print("%-12s %-28s FAILED: %r" % (args.workload, spec, e), flush=True)
        torch.cuda.empty_cache()
    engine.set_dfs("off")


if __name__ == "__main__":
    main()
