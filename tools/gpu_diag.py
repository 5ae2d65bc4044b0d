"""On-GPU diagnostic battery: every kernel against the CPU oracle, one subprocess per group so that a trapped
kernel cannot poison the others.  Usage (on the GPU box):

    python tools/gpu_diag.py            # all groups, prints a table, writes gpurun_out/diag.json
    python tools/gpu_diag.py --group gemm

Test infrastructure only (imports oracle/).
"""
import argparse
import json
import os
import subprocess
import sys
import time
import traceback
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["aux", "simt", "gemm", "conv", "stem", "attention", "models", "models_simt", "trn", "nlforced"]


def rel_err(got, ref):
    ref = ref.double()
    got = got.double().cpu()
    denom = max(ref.abs().max().item(), 1e-12)
    return (got - ref).abs().max().item() / denom


def run_group(group):
    import torch
    import torch.nn.functional as F
    import pretorched_x_b200 as P
    from pretorched_x_b200 import ops, engine
    from oracle import functional as OF

    dev = torch.device("cuda:0")
    results = []

    def record(name, err, tol, extra=""):
        ok = bool(err <= tol)
        results.append(dict(group=group, name=name, err=float(err), tol=tol, ok=ok, extra=extra))
        print("%-9s %-58s err %.3e tol %.1e %s %s" % (group, name, err, tol, "ok" if ok else "FAIL", extra), flush=True)

    def h(x):  # fp16 rounding on CPU (what the GPU path sees as input)
        return x.half().float()

    def act_from(x_ncdhw, pitch=None):
        return ops.from_ncdhw(x_ncdhw.to(dev), pitch=pitch)

    def conv_case(name, N, Cin, T, H, W, K, k, s, p, res=False, relu=True, bias=False, bn=True, simt=False, tol=2e-3):
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 10000)
        x = h(torch.randn(N, Cin, T, H, W, generator=g))
        conv = torch.nn.Conv3d(Cin, K, k, stride=s, padding=p, bias=bias)
        with torch.no_grad():
            conv.weight.copy_(h(torch.randn(conv.weight.shape, generator=g) * (1.0 / (Cin * k[0] * k[1] * k[2]) ** 0.5)))
            if bias:
                conv.bias.copy_(torch.randn(K, generator=g))
        bnm = None
        if bn:
            bnm = torch.nn.BatchNorm3d(K)
            OF.randomize_bn_(bnm, 7)
            bnm.eval()
        with torch.no_grad():
            y = conv(x)
            if bnm is not None:
                y = bnm(y)
            r = None
            if res:
                r = h(torch.randn(y.shape, generator=g))
                y = y + r
            if relu:
                y = F.relu(y)
        conv_d, bn_d = conv.to(dev), (bnm.to(dev) if bnm is not None else None)
        a = act_from(x)
        ra = act_from(r, pitch=ops._round_up(K, 8)) if res else None
        out = engine.conv_bn_act(conv_d, bn_d, a, residual=ra, relu=relu, simt=simt)
        torch.cuda.synchronize()
        got = ops.to_ncdhw(out)
        pad = out.data[:, out.C:]
        extra = "" if pad.numel() == 0 or float(pad.abs().max()) == 0 else "PAD-NONZERO"
        record(name, rel_err(got, y), tol, extra)

    if group == "aux":
        g = torch.Generator().manual_seed(1)
        x = torch.randn(2, 24, 5, 12, 10, generator=g)
        a = act_from(x)
        record("layout roundtrip C=24", rel_err(ops.to_ncdhw(a), h(x)), 1e-6)
        x3 = torch.randn(2, 3, 4, 8, 8, generator=g)
        a3 = act_from(x3)
        assert a3.ld == 4
        record("layout NDHWC4 (stem input)", rel_err(ops.to_ncdhw(a3), h(x3)), 1e-6)
        mp = ops.maxpool3d(a, (3, 3, 3), (2, 2, 2), (1, 1, 1))
        record("maxpool3d k3 s2 p1", rel_err(ops.to_ncdhw(mp), F.max_pool3d(h(x), 3, 2, 1)), 1e-6)
        mp2 = ops.maxpool3d(a, (1, 3, 3), (1, 2, 2), (0, 1, 1))
        record("maxpool (1,3,3)", rel_err(ops.to_ncdhw(mp2), F.max_pool3d(h(x), (1, 3, 3), (1, 2, 2), (0, 1, 1))), 1e-6)
        ap = ops.avgpool_global(a)
        record("avgpool global", rel_err(ap[:, :24].float(), h(x).mean(dim=(2, 3, 4))), 1e-3)
        sa = ops.shortcut_a(a, 2, 40)
        record("shortcut A s2 24->40", rel_err(ops.to_ncdhw(sa), OF.shortcut_a(h(x), 40, 2)), 1e-6)
        xr = torch.randn(7, 50, generator=g)
        c = ops.cast_rows(xr.to(dev), relu=True)
        record("cast+relu rows", rel_err(c[:, :50].float(), h(F.relu(xr))), 1e-6, "pad %.1f" % float(c[:, 50:].abs().max()))
        xf = torch.randn(3, 8, 16, generator=g).half()
        idx = torch.tensor([1, 4, 6], dtype=torch.int32)
        gf = ops.gather_frames(xf.to(dev), idx.to(dev))
        record("gather frames", rel_err(gf.float(), xf[:, [1, 4, 6], :].reshape(3, -1).float()), 1e-6)

    elif group == "simt":
        conv_case("simt 3x3x3 s1 C16->24", 1, 16, 4, 6, 6, 24, (3, 3, 3), (1, 1, 1), (1, 1, 1), res=True, simt=True)
        conv_case("simt 1x1x1 s2 C16->40", 1, 16, 4, 6, 6, 40, (1, 1, 1), (2, 2, 2), (0, 0, 0), relu=False, simt=True)
        conv_case("simt stem 7x7x7 C3->16", 1, 3, 4, 16, 16, 16, (7, 7, 7), (1, 2, 2), (3, 3, 3), simt=True)
        conv_case("simt stem (1,7,7) C3->24", 1, 3, 4, 16, 16, 24, (1, 7, 7), (1, 2, 2), (0, 3, 3), simt=True)

    elif group == "gemm":
        def gemm_case(name, M, N, Kd, res=False, relu=False, per_row=False, f32=False, acc=False, tol=2e-3):
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 10000)
            A = h(torch.randn(M, Kd, generator=g)); B = h(torch.randn(N, Kd, generator=g) / Kd ** 0.5)
            nsc = M if per_row else N
            sc = torch.rand(nsc, generator=g) + 0.5; sh = torch.randn(nsc, generator=g)
            D = A @ B.t()
            D = D * (sc.view(-1, 1) if per_row else sc.view(1, -1)) + (sh.view(-1, 1) if per_row else sh.view(1, -1))
            R = None
            if res:
                R = h(torch.randn(M, N, generator=g)); D = D + R
            if relu:
                D = F.relu(D)
            Kp, Np = ops._round_up(Kd, 8), ops._round_up(N, 8)
            Ad = torch.zeros(M, Kp, dtype=torch.float16, device=dev); Ad[:, :Kd] = A.half().to(dev)
            Bd = torch.zeros(N, Kp, dtype=torch.float16, device=dev); Bd[:, :Kd] = B.half().to(dev)
            Rd = None
            if res:
                Rd = torch.zeros(M, Np, dtype=torch.float16, device=dev); Rd[:, :N] = R.half().to(dev)
            out = None
            base = None
            if acc:
                base = torch.randn(M, N, generator=g)
                out = base.clone().to(dev)
                D = D + base
            got = ops.gemm(Ad, Bd, sc.to(dev), sh.to(dev), M, N, Kd, residual=Rd, relu=relu, per_row=per_row,
                           out_f32=f32, out=out, accumulate=acc)
            torch.cuda.synchronize()
            record(name, rel_err(got[:, :N].float(), D), tol)

        gemm_case("gemm 128x64x64 (1 tile, 1 kblock)", 128, 64, 64)
        gemm_case("gemm 128x64x256", 128, 64, 256)
        gemm_case("gemm 128x128x64 (BN=128)", 128, 128, 64)
        gemm_case("gemm 256x128x128 relu", 256, 128, 128, relu=True)
        gemm_case("gemm 300x200x192 res+relu (tails)", 300, 200, 192, res=True, relu=True)
        gemm_case("gemm 1000x512x512 res", 1000, 512, 512, res=True)
        gemm_case("gemm 64x400x2048 f32 (head)", 64, 400, 2048, f32=True)
        gemm_case("gemm 3x339x2048 f32 (tiny M)", 3, 339, 2048, f32=True)
        gemm_case("gemm 256x136x96 per_row (swap-AB)", 256, 136, 96, per_row=True)
        gemm_case("gemm 20x64x128 f32 accumulate", 20, 64, 128, f32=True, acc=True)
        gemm_case("gemm 5x1024x16384 relu (TRN fc1)", 5, 512, 16384, relu=True)
        gemm_case("gemm 4096x256x64 (layer1 conv3 shape)", 4096, 256, 64, relu=True)

    elif group == "conv":
        conv_case("1x1x1 s1 64->64 (TMA A)", 2, 64, 4, 8, 8, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0))
        conv_case("1x1x1 s1 64->256 +res", 2, 64, 4, 8, 8, 256, (1, 1, 1), (1, 1, 1), (0, 0, 0), res=True)
        conv_case("1x1x1 s1 256->64", 1, 256, 4, 8, 8, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0))
        conv_case("3x3x3 s1 64->64 (gather)", 1, 64, 4, 8, 8, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        conv_case("3x3x3 s1 64->64 M-tail", 1, 64, 3, 7, 5, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        conv_case("3x3x3 s2 128->128", 1, 128, 4, 8, 8, 128, (3, 3, 3), (2, 2, 2), (1, 1, 1))
        conv_case("3x3x3 s1 128->128 +res", 2, 128, 2, 6, 6, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), res=True)
        conv_case("1x1x1 s2 256->512 (shortcut B)", 1, 256, 4, 8, 8, 512, (1, 1, 1), (2, 2, 2), (0, 0, 0), relu=False)
        conv_case("(1,3,3) 64->144 (r2p1d spatial)", 1, 64, 4, 8, 8, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1))
        conv_case("(3,1,1) 144->64 (r2p1d temporal)", 1, 144, 4, 8, 8, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0))
        conv_case("(7,1,1) 110->64 (r2p1d stem temporal)", 1, 110, 8, 8, 8, 64, (7, 1, 1), (1, 1, 1), (3, 0, 0))
        conv_case("(1,3,3) s(1,2,2) 64->230", 1, 64, 4, 8, 8, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1))
        conv_case("(3,1,1) s(2,1,1) 230->128", 1, 230, 4, 4, 4, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0))
        conv_case("1x1x1 bias no-bn 128->64 (theta)", 1, 128, 2, 4, 4, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), bias=True, bn=False, relu=False)
        conv_case("3x3x3 s1 512->512 T=1 (layer4)", 2, 512, 1, 7, 7, 512, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        conv_case("3x3x3 s1 64->64 big M", 2, 64, 8, 28, 28, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        conv_case("slab 3x3x3 64->64 56x56 T=4 (layer1)", 2, 64, 4, 56, 56, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        conv_case("slab 3x3x3 128->128 28x28 T=3 +res", 2, 128, 3, 28, 28, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1), res=True)
        conv_case("slab 3x3x3 256->256 14x14 T=2", 3, 256, 2, 14, 14, 256, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        conv_case("slab 3x3x3 512->512 7x7 T=1", 3, 512, 1, 7, 7, 512, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        conv_case("slab 3x3x3 64->200 odd sizes 13x11 T=5", 1, 64, 5, 13, 11, 200, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        conv_case("slab (1,3,3) 64->144 28x28", 1, 64, 4, 28, 28, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1))
        conv_case("slab (3,1,1) 144->64 28x28", 1, 144, 6, 28, 28, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), res=True)
        conv_case("slab (7,1,1) 110->64 56x56", 1, 110, 8, 56, 56, 64, (7, 1, 1), (1, 1, 1), (3, 0, 0))
        conv_case("slab 3x3 2-D 64->64 56x56 (resnet18)", 2, 64, 1, 56, 56, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), res=True)
        conv_case("slab 3x3x3 64->64 112x112 T=2 (wide)", 1, 64, 2, 112, 112, 64, (3, 3, 3), (1, 1, 1), (1, 1, 1))

    elif group == "stem":
        conv_case("stem 7x7x7 s(1,2,2) 3->64", 1, 3, 4, 32, 32, 64, (7, 7, 7), (1, 2, 2), (3, 3, 3))
        conv_case("stem 7x7x7 3->64 B=2 T=8 64x64", 2, 3, 8, 64, 64, 64, (7, 7, 7), (1, 2, 2), (3, 3, 3))
        conv_case("stem (1,7,7) 3->110 (r2p1d)", 1, 3, 4, 32, 32, 110, (1, 7, 7), (1, 2, 2), (0, 3, 3))
        conv_case("stem 2-D 7x7 s2 (T=1) 3->64", 2, 3, 1, 32, 32, 64, (1, 7, 7), (1, 2, 2), (0, 3, 3))

    elif group == "attention":
        def att_case(name, B, Npos, d, dv, scale=0.3, tol=4e-3):
            g = torch.Generator().manual_seed(zlib.crc32(name.encode()) % 10000)
            q = h(torch.randn(B, Npos, d, generator=g) * scale)
            k = h(torch.randn(B, Npos, d, generator=g) * scale)
            v = h(torch.randn(B, Npos, dv, generator=g))
            ref = torch.softmax(q @ k.transpose(1, 2), dim=-1) @ v
            qkv = torch.cat([q, k, v], dim=2).reshape(B * Npos, 2 * d + dv).half().to(dev).contiguous()
            o = ops.nonlocal_attention(qkv, d, dv, B, Npos)
            torch.cuda.synchronize()
            record(name, rel_err(o[:, :dv].float().view(B, Npos, dv), ref), tol)

        att_case("att B1 N128 d64 dv64", 1, 128, 64, 64)
        att_case("att B2 N256 d64 dv64", 2, 256, 64, 64)
        att_case("att B2 N200 d128 dv128 (tails)", 2, 200, 128, 128)
        att_case("att B1 N576 d256 dv256", 1, 576, 256, 256)
        att_case("att B2 N72 d512 dv512 (dv split)", 2, 72, 512, 512)
        att_case("att B1 N784 d512 dv512", 1, 784, 512, 512)
        att_case("att peaked logits (scale 3)", 1, 256, 64, 64, scale=3.0, tol=2e-2)
        for cfg in [(1, 98, 64, 64), (2, 128, 128, 128), (1, 98, 128, 128), (2, 100, 128, 128), (2, 98, 128, 128)]:
            att_case("att sweep B%d N%d d%d dv%d" % cfg, *cfg)

    elif group in ("models", "models_simt"):
        import glob
        simt = group == "models_simt"
        for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.pt"))):
            fx = torch.load(path, weights_only=False)
            if fx["kind"] != "model":
                continue
            name = os.path.basename(path)[:-3]
            if simt and "nonlocal" in name:
                continue
            torch.manual_seed(fx["seeds"]["init"])
            arch = fx["arch"]
            m = getattr(P, arch)(**fx["kwargs"]) if arch.startswith("r2") else getattr(P, arch)(pretrained=None, **fx["kwargs"])
            OF.randomize_bn_(m, fx["seeds"]["bn"])
            if fx.get("nl_factors"):
                OF.apply_nonlocal_factors_(m, fx["nl_factors"])
            m = m.eval().to(dev)
            x = OF.seeded_input(fx["input_shape"], fx["seeds"]["input"]).to(dev)
            t0 = time.time()
            with torch.no_grad():
                a = engine.run_stem(m, x, simt=simt)
                stage_out = {"maxpool": a}
                for ln in ("layer1", "layer2", "layer3", "layer4"):
                    for blk in getattr(m, ln):
                        a = engine.run_block(blk, a, simt=simt)
                    stage_out[ln] = a
                logits = m.logits(a)
            torch.cuda.synchronize()
            for sn, ref in fx["stages"].items():
                if sn == "logits":
                    continue
                t = ops.to_ncdhw(stage_out[sn])
                if t.dim() == 5 and len(ref["shape"]) == 4:
                    t = t.squeeze(2)
                samp = t.reshape(-1)[::ref["step"]][:ref["sample"].numel()].cpu()
                err = (samp.double() - ref["sample"].double()).abs().max().item() / max(ref["absmax"], 1e-12)
                record("%s/%s" % (name, sn), err, 1e-2)
            err = rel_err(logits, fx["logits"])
            agree = bool((logits.argmax(1).cpu() == fx["logits"].argmax(1)).all())
            record("%s/logits" % name, err, 5e-3 if "nonlocal" not in name else 2e-2,
                   "argmax %s %.2fs" % ("agree" if agree else "DIFFER", time.time() - t0))

    elif group == "trn":
        import numpy as np
        gd = os.path.join(ROOT, "tests", "golden")
        for name in ("relation_small", "relation_htrn"):
            fx = torch.load(os.path.join(gd, name + ".pt"), weights_only=False)
            torch.manual_seed(fx["seeds"]["init"])
            r = P.Relation(fx["T"], fx["F"], fx["out"], bottleneck_dim=fx["bottleneck"]).to(dev).eval()
            x = OF.seeded_input((fx["B"], fx["T"], fx["F"]), fx["seeds"]["input"]).to(dev)
            with torch.no_grad():
                y = r(x)
            record(name, rel_err(y, fx["output"]), 5e-3)
        fx = torch.load(os.path.join(gd, "msrelation_small.pt"), weights_only=False)
        torch.manual_seed(fx["seeds"]["init"])
        r = P.MultiScaleRelation(fx["T"], fx["F"], fx["out"], bottleneck_dim=fx["bottleneck"]).to(dev).eval()
        x = OF.seeded_input((fx["B"], fx["T"], fx["F"]), fx["seeds"]["input"]).to(dev)
        np.random.seed(fx["np_seed"])
        with torch.no_grad():
            y = r(x)
        record("msrelation_small", rel_err(y, fx["output"]), 5e-3)

    elif group == "nlforced":
        # teacher-forced per-block parity of the non-local net: feed OUR block input to the CPU oracle block, so
        # accumulated upstream error is excluded and only the block's own arithmetic is compared.
        fx = torch.load(os.path.join(ROOT, "tests", "golden", "nonlocalresnet3d50_b1_t16_96.pt"), weights_only=False)
        torch.manual_seed(fx["seeds"]["init"])
        m = P.nonlocalresnet3d50(pretrained=None)
        OF.randomize_bn_(m, fx["seeds"]["bn"])
        sd = {k: v.clone() for k, v in m.state_dict().items()}
        m = m.eval().to(dev)
        x = OF.seeded_input(fx["input_shape"], fx["seeds"]["input"]).to(dev)
        nlpos = OF.nonlocal_positions([3, 4, 6, 3], [0, 2, 3, 0])
        with torch.no_grad():
            a = engine.run_stem(m, x)
            inplanes = 64
            for li, ln in enumerate(("layer1", "layer2", "layer3", "layer4")):
                planes = (64, 128, 256, 512)[li]
                for bi, blk in enumerate(getattr(m, ln)):
                    xin = ops.to_ncdhw(a).cpu()
                    stride = 2 if (li > 0 and bi == 0) else 1
                    has_ds = bi == 0 and (stride != 1 or inplanes != planes * 4)
                    pfx = "%s.%d" % (ln, bi)
                    ref_b = OF.bottleneck(xin, sd, pfx, "resnet3d", "A", planes, stride, has_ds)
                    mid = engine.run_bottleneck(blk, a)
                    e_b = rel_err(ops.to_ncdhw(mid), ref_b)
                    record("%s bottleneck" % pfx, e_b, 3e-3)
                    if bi in nlpos[li]:
                        xin2 = ops.to_ncdhw(mid).cpu()
                        ref_n = OF.nonlocal_block(xin2, sd, pfx + ".nonlocalblock")
                        out = engine.run_nonlocal(blk.nonlocalblock, mid)
                        got = ops.to_ncdhw(out).cpu()
                        d = (got.double() - ref_n.double()).abs() / ref_n.abs().max().item()
                        # per-position error: a flipped arg-max shows up as whole positions being off
                        dpos = d.amax(dim=1).reshape(-1)
                        frac_bad = float((dpos > 1e-2).double().mean())
                        # logit scale of this block
                        th = OF._conv(xin2, sd, pfx + ".nonlocalblock.theta").flatten(2)
                        ph = OF._conv(xin2, sd, pfx + ".nonlocalblock.phi").flatten(2)
                        f = th.transpose(1, 2) @ ph
                        top2 = f.topk(2, dim=-1).values
                        gap = (top2[..., 0] - top2[..., 1])
                        record("%s nonlocal" % pfx, d.max().item(), 1e-2,
                               "median %.2e, positions>1e-2: %.1f%%, |f|max %.2e, top2 gap min %.2e med %.2e"
                               % (d.median().item(), 100 * frac_bad, f.abs().max().item(), gap.min().item(), gap.median().item()))
                        a = out
                    else:
                        a = mid
                    inplanes = planes * 4

    return results


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", default=None)
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    if args.group:
        try:
            res = run_group(args.group)
        except Exception:
            traceback.print_exc()
            res = [dict(group=args.group, name="EXCEPTION", err=float("inf"), tol=0, ok=False,
                        extra=traceback.format_exc().splitlines()[-1][:200])]
        if args.json:
            with open(args.json, "w") as fh:
                json.dump(res, fh)
        sys.exit(0 if all(r["ok"] for r in res) else 1)

    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    allres = []
    for gname in GROUPS:
        jpath = os.path.join(out_dir, "diag_%s.json" % gname)
        if os.path.exists(jpath):
            os.remove(jpath)
        try:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--group", gname, "--json", jpath], timeout=420)
        except subprocess.TimeoutExpired:
            print("%-9s TIMEOUT" % gname, flush=True)
        if os.path.exists(jpath):
            allres += json.load(open(jpath))
        else:
            allres.append(dict(group=gname, name="NO RESULT (crash/timeout)", err=float("inf"), tol=0, ok=False, extra=""))
    with open(os.path.join(out_dir, "diag.json"), "w") as fh:
        json.dump(allres, fh, indent=1)
    nfail = sum(1 for r in allres if not r["ok"])
    print("\n== diag: %d checks, %d failed ==" % (len(allres), nfail))
    for r in allres:
        if not r["ok"]:
            print("FAIL %-9s %-58s err %.3e %s" % (r["group"], r["name"], r["err"], r["extra"]))


if __name__ == "__main__":
    main()
