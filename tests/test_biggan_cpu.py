"""CPU: the BigGAN-deep generator container and its restatement (oracle/biggan.py -- PARITY UNPINNED: the reference tree
has no GAN code).  Checks the published state_dict layout, the spectral-norm arithmetic,
the algorithmic MAC count used by the bench, that the restatement still reproduces its committed fixtures, and that the
product path refuses to run without a GPU."""
import glob
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import biggan as OB
from oracle import functional as OF
import pretorched_x_b200 as P

FIX = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "biggan_*.pt")))


def test_state_dict_layout_of_the_published_generator():
    torch.manual_seed(0)
    m = P.biggan_deep(256, G_ch=16, n_classes=10)
    sd = m.state_dict()
    ch = 16
    assert sd["shared.weight"].shape == (10, 128)
    assert sd["linear.weight"].shape == (16 * ch * 16, 256) and sd["linear.u0"].shape == (1, 16 * ch * 16)
    assert sd["linear.sv0"].shape == (1,)
    # stage 3: 8ch -> 4ch, followed by attention on 4ch channels
    assert sd["blocks.3.1.conv1.weight"].shape == (2 * ch, 8 * ch, 1, 1)
    assert sd["blocks.3.1.conv2.weight"].shape == (2 * ch, 2 * ch, 3, 3)
    assert sd["blocks.3.1.conv4.weight"].shape == (4 * ch, 2 * ch, 1, 1) and sd["blocks.3.1.conv4.bias"].shape == (4 * ch,)
    assert sd["blocks.3.1.bn1.gain.weight"].shape == (8 * ch, 256) and "blocks.3.1.bn1.gain.bias" not in sd
    assert sd["blocks.3.1.bn1.bias.u0"].shape == (1, 8 * ch)
    assert sd["blocks.3.1.bn4.stored_var"].shape == (2 * ch,)
    assert sd["blocks.3.2.theta.weight"].shape == (4 * ch // 8, 4 * ch, 1, 1) and "blocks.3.2.theta.bias" not in sd
    assert sd["blocks.3.2.g.weight"].shape == (2 * ch, 4 * ch, 1, 1) and sd["blocks.3.2.o.weight"].shape == (4 * ch, 2 * ch, 1, 1)
    assert sd["blocks.3.2.gamma"].shape == ()
    assert "blocks.2.2.theta.weight" not in sd and "blocks.4.2.theta.weight" not in sd      # attention at 64x64 only
    assert sd["output_layer.0.gain"].shape == (ch,) and sd["output_layer.2.weight"].shape == (3, ch, 3, 3)
    # parameter order inside a spectrally normalised layer: weight, bias, then the u / sv buffers
    keys = [k for k in sd if k.startswith("blocks.0.0.conv1.")]
    assert keys == ["blocks.0.0.conv1.weight", "blocks.0.0.conv1.bias", "blocks.0.0.conv1.u0", "blocks.0.0.conv1.sv0"]
    assert len(m.blocks) == 6 and [len(s) for s in m.blocks] == [2, 2, 2, 3, 2, 2]


def test_full_size_parameter_count():
    m = P.biggan_deep256()
    n = sum(p.numel() for p in m.parameters())
    assert 55.0e6 < n < 56.5e6           # 55.7 M for ch = 128, 1000 classes
    assert m.arch["in_channels"] == [2048, 2048, 1024, 1024, 512, 256] and m.arch["out_channels"] == [2048, 1024, 1024, 512, 256, 128]


def test_spectral_norm_sigma():
    g = torch.Generator().manual_seed(3)
    w = torch.randn(24, 40, generator=g)
    u = torch.randn(1, 24, generator=g)
    s = OB.sn_sigma(w, u)
    top = torch.linalg.svdvals(w)[0]
    assert 0 < s <= top * (1 + 1e-6)
    # iterating the same step converges to the largest singular value
    for _ in range(200):
        v = F.normalize(u @ w)
        u = F.normalize(v @ w.t())
    assert abs(OB.sn_sigma(w, u) - top) <= 1e-4 * top
    # orthogonal rows / columns: sigma == 1 from any start vector
    q = torch.nn.init.orthogonal_(torch.empty(16, 64))
    assert abs(OB.sn_sigma(q, torch.randn(1, 16)) - 1.0) < 1e-5


def test_spectral_norm_sigma_equals_torch_spectral_norm_power_step():
    """The restatement's sigma (one power-iteration step from the stored u0, eval semantics of the published generator) is the
    sigma torch.nn.utils.spectral_norm -- an independent, third-party implementation of the same estimator -- computes in its first
    training-mode forward from the same start vector, for a Linear and for a 3x3 convolution."""
    import torch.nn as nn
    torch.manual_seed(11)
    for mod, x in ((nn.Linear(40, 24, bias=False), torch.zeros(1, 40)), (nn.Conv2d(6, 10, 3, bias=False), torch.zeros(1, 6, 5, 5))):
        sn = torch.nn.utils.spectral_norm(mod, n_power_iterations=1)
        u0 = sn.weight_u.detach().clone()
        w = sn.weight_orig.detach().clone()
        sn.train()
        sn(x)                                                  # one power iteration from u0; module.weight = weight_orig / sigma
        sigma_torch = (w.norm() / sn.weight.detach().norm()).item()
        ours = OB.sn_sigma(w, u0.view(1, -1)).item()
        assert abs(ours - sigma_torch) <= 1e-5 * sigma_torch, (ours, sigma_torch)
        assert torch.allclose(w / ours, sn.weight.detach(), rtol=1e-5, atol=1e-7)


def test_mac_count_matches_hooked_restatement(monkeypatch):
    res, ch, ncls, B = 128, 16, 10, 1
    _, sd, z, labels = OB.build_case(P.biggan_deep, res, ch, ncls, B)
    macs = [0]
    conv2d, linear, bmm = F.conv2d, F.linear, torch.bmm

    def c2(x, w, b=None, stride=1, padding=0):
        y = conv2d(x, w, b, stride, padding)
        macs[0] += y.numel() * w.shape[1] * w.shape[2] * w.shape[3]
        return y

    def lin(x, w, b=None):
        macs[0] += x.shape[0] * w.numel()
        return linear(x, w, b)

    def bm(a, b):
        macs[0] += a.shape[0] * a.shape[1] * a.shape[2] * b.shape[2]
        return bmm(a, b)

    monkeypatch.setattr(OB.F, "conv2d", c2)
    monkeypatch.setattr(OB.F, "linear", lin)
    monkeypatch.setattr(OB.torch, "bmm", bm)
    with torch.no_grad():
        OB.generator_forward(z, labels, sd, res, ch)
    assert macs[0] == OB.mac_count(res, ch)
    assert abs(OB.mac_count(256, 128) / 1e9 - 29.4) < 0.1          # BigGAN-deep-256: 29.4 GMAC = 58.8 GFLOP per image


@pytest.mark.parametrize("path", FIX, ids=[os.path.basename(p)[:-3] for p in FIX])
def test_restatement_reproduces_its_fixture(path):
    fx = torch.load(path, weights_only=False)
    model, sd, z, labels = OB.build_case(P.biggan_deep, fx["resolution"], fx["ch"], fx["n_classes"], fx["batch"],
                                         fx["seeds"]["init"], fx["seeds"]["input"], fx["init"])
    assert list(sd) == fx["keys"] and len(sd) == fx["n_state"]
    assert OF.digests_match(OF.state_digest(sd), fx["weight_digest"], rtol=1e-4)
    stages = {}
    with torch.no_grad():
        img = OB.generator_forward(z, labels, sd, fx["resolution"], fx["ch"], stages=stages)
    stages["image"] = img
    for name, ref in list(fx["stages"].items()) + [("image", fx["image"])]:
        got = stages[name]
        assert tuple(got.shape) == ref["shape"]
        samp = got.reshape(-1)[::ref["step"]][:ref["sample"].numel()]
        assert (samp - ref["sample"]).abs().max().item() <= 2e-4 * max(ref["absmax"], 1e-6), name


def test_fixtures_present():
    assert len(FIX) >= 2


def test_no_cpu_path():
    m = P.biggan_deep(128, G_ch=16, n_classes=10)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(1, 128), torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError, match="inference-only"):
        m.train()(torch.randn(1, 128), torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError):
        m.blocks[0][0].conv1(torch.randn(1, 256, 4, 4))


def test_stacked_ccbn_gemm_algebra(monkeypatch):
    """Host logic of biggan_engine._pack, checked without a GPU: every conditional BatchNorm of the generator is an affine
    map per (sample, channel) that ONE GEMM of the conditioning vector against the stacked, pre-folded gain / bias matrices
    produces -- incl. the conv bias folded into the mean when the ccbn rides in a convolution's epilogue, and the fp16
    hi/lo operand split ([y_hi | y_lo | y_hi] . [W_hi | W_hi | W_lo]^T) that keeps the gains at ~fp32 accuracy."""
    from pretorched_x_b200 import biggan_engine as BE

    class FakePacked:                                   # ops.PackedConv needs the CUDA library; the algebra under test does not
        def __init__(self, w, bias, bn, stride, padding, in_pitch=None, upsample=False):
            self.K, self.scale, self.shift = w.shape[0], None, None
    monkeypatch.setattr(BE.ops, "PackedConv", FakePacked)

    res, ch, ncls, B = 128, 16, 10, 3
    model, sd, z, labels = OB.build_case(P.biggan_deep, res, ch, ncls, B)
    with torch.no_grad():
        pk = BE._pack(model, torch.device("cpu"))
        y = OB.condition(z, labels, sd)                                      # [B][256] fp32
        y_hi = y.half()
        y_lo = (y - y_hi.float()).half()
        y3 = torch.cat([y_hi, y_lo, y_hi], 1).double()                       # what b2_embed_concat(split=1) writes
        aff = y3 @ pk.cond_w.double().t() * pk.cond_scale.double() + pk.cond_shift.double()      # the GEMM + its epilogue
        assert pk.cond_w.shape == (pk.ncols, 3 * 256) and pk.ncols == 2 * sum(
            b.bn1.output_size + 3 * b.bn2.output_size for st in model.blocks for b in st if hasattr(b, "conv4"))
        g = torch.Generator().manual_seed(4)
        worst = 0.0
        for st_i, stage in enumerate(model.blocks):
            for b_i, blk in enumerate(stage):
                if not hasattr(blk, "conv4"):
                    continue
                bp = pk.blocks[id(blk)]
                prev = [None, blk.conv1.bias if bp.fuse2 else None, blk.conv2.bias, blk.conv3.bias]
                for k, (bn, pb) in enumerate(zip((blk.bn1, blk.bn2, blk.bn3, blk.bn4), prev)):
                    o_s, o_t, C = bp.bn[k]
                    x = torch.randn(B, C, 3, 3, generator=g) * 2 + 1                 # "accumulator" (before the conv bias)
                    xin = x + (pb.detach().view(1, C, 1, 1) if pb is not None else 0)
                    want = OB.ccbn(xin, y, sd, "blocks.%d.%d.bn%d" % (st_i, b_i, k + 1)).double()
                    got = x.double() * aff[:, o_s:o_s + C].view(B, C, 1, 1) + aff[:, o_t:o_t + C].view(B, C, 1, 1)
                    worst = max(worst, ((got - want).abs().max() / want.abs().max()).item())
        assert worst <= 2e-5, worst                     # fp32 reference arithmetic is the floor; plain fp16 operands give ~1e-3
        # first linear: rows permuted so that the GEMM output is already channels-last [B*16][C0]
        C0 = model.arch["in_channels"][0]
        h = (y.half().double() @ pk.lin_w.double()[:, :256].t() + pk.lin_b.double()).view(B, 16, C0).permute(0, 2, 1).reshape(B, C0, 4, 4)
        want_h = torch.nn.functional.linear(y, OB.sn_weight(sd, "linear"), sd["linear.bias"]).view(B, C0, 4, 4)
        assert ((h - want_h.double()).abs().max() / want_h.abs().max()).item() <= 2e-3


def test_generator_dfs_plan_levels():
    """biggan_engine.dfs_plan: modules are grouped by OUTPUT resolution into a contiguous suffix of depth-first levels."""
    import pretorched_x_b200 as P
    from pretorched_x_b200 import biggan_engine as be
    m = P.biggan_deep(256, G_ch=16, n_classes=10).eval()
    mods = [(i, blk) for i, stage in enumerate(m.blocks) for blk in stage]
    assert len(mods) == 13                                    # 6 stages x 2 GBlocks + attention at 64x64
    try:
        be.set_dfs("off")
        assert be.dfs_plan(m, 256, mods) == []
        be.set_dfs("256:4")
        assert be.dfs_plan(m, 256, mods) == [(12, 13, 4)]
        be.set_dfs("64:16,128:8,256:4")
        assert be.dfs_plan(m, 256, mods) == [(7, 10, 16), (10, 12, 8), (12, 13, 4)]
        be.set_dfs("64:8,128:300,256:4")                      # a whole-batch level behind a chunked one stays in the suffix
        assert be.dfs_plan(m, 256, mods) == [(7, 10, 8), (10, 12, 256), (12, 13, 4)]
        be.set_dfs("128:300,256:4")                           # ... in front of it: part of the whole-batch prefix
        assert be.dfs_plan(m, 256, mods) == [(12, 13, 4)]
        be.set_dfs("256:4")
        assert be.dfs_plan(m, 4, mods) == []                  # batch no larger than the chunk
    finally:
        be.set_dfs("off")
