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
