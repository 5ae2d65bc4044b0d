"""CPU: host-side logic of the drop-in boundary -- factory registry, state_dict layout, reference quirks,
BN folding, sharding arithmetic, and loud failure without a GPU."""
import math

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import pretorched_x_b200 as P
from pretorched_x_b200 import ops, parallel
from pretorched_x_b200.models import r2plus1d, nonlocalnet
from oracle import functional as OF


def test_factory_registry_matches_reference_names():
    for name in ["resnet18", "resnet3d10", "resnet3d18", "resnet3d34", "resnet3d50", "resnet3d101", "resnet3d152",
                 "resnet3d200", "resneti3d50", "nonlocalresnet3d18", "nonlocalresnet3d34", "nonlocalresnet3d50",
                 "nonlocalresnet3d101"]:
        assert callable(P.__dict__[name]), name
    s = P.pretrained_settings["resnet3d50"]["kinetics-400"]
    assert s["num_classes"] == 400 and s["input_size"] == [3, 224, 224] and s["mean"] == [0.485, 0.456, 0.406]
    assert P.pretrained_settings["resnet3d50"]["moments"]["num_classes"] == 339
    assert "resnet3d50" in P.model_names and "resnet18" in P.model_names


def test_resnet3d50_state_dict_layout():
    m = P.resnet3d50(num_classes=400, pretrained=None)
    sd = m.state_dict()
    assert len(sd) == 320
    assert tuple(sd["conv1.weight"].shape) == (64, 3, 7, 7, 7)
    assert tuple(sd["layer2.0.downsample.0.weight"].shape) == (512, 256, 1, 1, 1)
    assert tuple(sd["last_linear.weight"].shape) == (400, 2048) and m.fc is None
    assert not any(k.startswith("fc.") for k in sd)


def test_zoo_checkpoints_with_fc_keys_load():
    m = P.resnet3d10(num_classes=7)
    sd = {k.replace("last_linear.", "fc."): v.clone() for k, v in m.state_dict().items()}
    sd["fc.bias"] += 1.0
    m2 = P.resnet3d10(num_classes=7)
    m2.load_state_dict(sd)
    assert torch.equal(m2.last_linear.bias, sd.get("fc.bias", m2.last_linear.bias))


def test_pretrained_asserts_num_classes_before_download():
    with pytest.raises(AssertionError):
        P.resnet3d50(num_classes=10, pretrained="kinetics-400")


def test_r2plus1d_layout_and_quirks():
    assert r2plus1d.intermediate_channels(3, 64, (7, 7, 7)) == 110
    assert r2plus1d.intermediate_channels(64, 64, (3, 3, 3)) == 144
    assert r2plus1d.intermediate_channels(64, 128, (1, 1, 1)) == 42
    P.resnet3d10()                                                       # would break the reference's R2Plus1D
    m = P.r2plus1d34(num_classes=400)
    sd = m.state_dict()
    assert len(sd) == 434
    assert tuple(sd["conv1.spatial_conv.weight"].shape) == (110, 3, 1, 7, 7)
    assert tuple(sd["conv1.temporal_conv.weight"].shape) == (64, 110, 7, 1, 1)
    assert tuple(sd["fc.weight"].shape) == (400, 512) and m.last_linear is m.fc
    m.last_linear = P.Identity()
    assert isinstance(m.fc, P.Identity) and "last_linear" not in dict(m.named_children())


def test_nonlocal_quirks():
    m = P.nonlocalresnet3d50(num_classes=10, pretrained=None)            # num_classes is swallowed upstream
    assert m.last_linear.out_features == 339
    sd = m.state_dict()
    assert not any("downsample" in k for k in sd)                        # type-A shortcuts, no parameters
    nl = [k.split(".nonlocalblock")[0] for k in sd if k.endswith("nonlocalblock.g.weight")]
    assert nl == ["layer2.0", "layer2.2", "layer3.0", "layer3.2", "layer3.4"]
    assert float(m.layer2[0].nonlocalblock.W[1].weight.abs().mean()) == 1.0   # init_weights overrides the zero init
    assert OF.nonlocal_positions([3, 4, 6, 3], [0, 2, 3, 0]) == [[], [0, 2], [0, 2, 4], []]
    cat = nonlocalnet.NonLocalBlock3D(64, mode="concatenation")               # nonlocalnet.py:117-121: 2d -> 1 projection, no bias
    assert tuple(cat.state_dict()["concat_project.0.weight"].shape) == (1, 64, 1, 1) and "concat_project.0.bias" not in cat.state_dict()
    sub = nonlocalnet.NonLocalBlock2D(64, mode="gaussian", sub_sample=True)     # reference layout with sub_sample
    assert set(sub.state_dict()) >= {"g.0.weight", "g.0.bias", "W.0.weight", "W.1.running_mean"}
    assert "theta.weight" not in sub.state_dict()


def test_trn_upstream_defects_are_explicit():
    with pytest.raises(NotImplementedError):
        P.HierarchicalRelation(8, 16, 8, relation_size=4)
    h = P.HierarchicalRelation(8, 16, 8, relation_size=1024)             # what TRN(consensus='HTRN') builds
    assert h.final_relation.bottleneck_dim == 512


def test_fold_affine_equals_eval_batchnorm():
    torch.manual_seed(0)
    bn = nn.BatchNorm3d(12)
    OF.randomize_bn_(bn, 5)
    bn.eval()
    bias = torch.randn(12)
    scale, shift = ops.fold_affine(12, bias, bn, torch.device("cpu"))
    x = torch.randn(2, 12, 3, 4, 5)
    want = bn(x + bias.view(1, -1, 1, 1, 1))
    got = x * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)
    bn.train()
    with pytest.raises(RuntimeError):
        ops.fold_affine(12, None, bn, torch.device("cpu"))


def test_forward_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = P.resnet3d10(num_classes=5).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.randn(1, 3, 4, 32, 32))
    m.train()
    with pytest.raises(RuntimeError, match="inference-only"):
        m(torch.randn(1, 3, 4, 32, 32))


def test_shard_bounds_cover_batch_exactly():
    for total in (1, 7, 16, 32, 33, 64):
        for world in (1, 2, 4, 8):
            spans = [parallel.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
