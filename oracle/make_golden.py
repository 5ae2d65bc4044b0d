"""Generate tests/golden/*.pt from the UNMODIFIED reference (run in the build container only):

    python -m oracle.make_golden

For every case the script (1) builds the reference model from /root/reference under a fixed seed, (2) gives its
BatchNorms non-trivial statistics (oracle.functional.randomize_bn_), (3) runs the reference's own CPU fp32
forward with per-stage forward hooks -- the pattern of the reference's only value-level check,
pretorched/models/fbresnet/resnet152_load.py:251-270 -- (4) asserts that oracle/functional.py reproduces every
stage bit-for-bit, and (5) stores logits, strided stage samples and a per-tensor weight digest.  The fixtures let
the GPU box (which has no /root/reference) verify that pretorched_x_b200's seeded init equals the reference's
and that its CUDA forward matches the reference's outputs.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import functional as OF            # noqa: E402
from oracle import reference_loader as RL      # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
SAMPLE = 4096

# name -> (arch, factory kwargs, input shape)
MODEL_CASES = {
    "resnet3d50_b2_t8_64": ("resnet3d50", dict(num_classes=400), (2, 3, 8, 64, 64)),
    "resnet3d18_b1_t8_64": ("resnet3d18", dict(num_classes=400), (1, 3, 8, 64, 64)),
    "r2plus1d34_b1_t8_64": ("r2plus1d34", dict(num_classes=400), (1, 3, 8, 64, 64)),
    "nonlocalresnet3d50_b1_t16_96": ("nonlocalresnet3d50", dict(), (1, 3, 16, 96, 96)),
    "resnet18_b2_64": ("resnet18", dict(num_classes=1000), (2, 3, 64, 64)),
    # same net with theta/phi rescaled into a trained-like logit regime (oracle.functional.calibrate_nonlocal_)
    "nonlocalresnet3d50_tamed_b1_t16_96": ("nonlocalresnet3d50", dict(), (1, 3, 16, 96, 96)),
}
SEED_INIT, SEED_BN, SEED_INPUT = 0, 1, 2


def sample_of(t):
    flat = t.detach().reshape(-1)
    step = max(1, flat.numel() // SAMPLE)
    return flat[::step][:SAMPLE].clone(), step


def summarize(t):
    s, step = sample_of(t)
    return dict(shape=tuple(t.shape), mean=float(t.double().mean()), std=float(t.double().std()),
                absmax=float(t.abs().max()), sample=s, step=step)


def run_model_case(name, arch, kwargs, shape):
    RL.load()
    if arch.startswith("r2plus1d"):
        RL.load_r2plus1d()
    torch.manual_seed(SEED_INIT)
    ref = RL.build(arch, **kwargs)
    OF.randomize_bn_(ref, SEED_BN)
    ref.eval()
    x = OF.seeded_input(shape, SEED_INPUT)
    nl_factors = OF.calibrate_nonlocal_(ref, x) if "tamed" in name else None

    hooked = {}
    handles = []
    for stage in ("maxpool", "layer1", "layer2", "layer3", "layer4"):
        handles.append(getattr(ref, stage).register_forward_hook(
            lambda m, i, o, stage=stage: hooked.__setitem__(stage, o.detach().clone())))
    with torch.no_grad():
        if arch.startswith("r2plus1d"):
            # R2Plus1D inherits ResNet3D.forward, which modify_resnets may have patched at class level to need
            # `last_linear`.  The unpatched body is conv1..layer4 -> avgpool -> fc:
            feat = ref.layer4(ref.layer3(ref.layer2(ref.layer1(ref.maxpool(ref.relu(ref.bn1(ref.conv1(x))))))))
            logits = ref.fc(ref.avgpool(feat).view(feat.size(0), -1))
        else:
            logits = ref(x)
    for h in handles:
        h.remove()
    hooked["logits"] = logits

    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    stages = {}
    with torch.no_grad():
        out = OF.forward(x, sd, arch, stages)
    for k, v in hooked.items():
        assert torch.equal(stages[k], v), "oracle restatement differs from the reference at %s/%s" % (name, k)
    assert torch.equal(out, logits)

    fixture = dict(
        kind="model", arch=arch, kwargs=kwargs, input_shape=tuple(shape), nl_factors=nl_factors,
        seeds=dict(init=SEED_INIT, bn=SEED_BN, input=SEED_INPUT),
        logits=logits.clone(), stages={k: summarize(v) for k, v in hooked.items()},
        weight_digest=OF.state_digest(sd), n_state=len(sd),
        torch_version=torch.__version__,
    )
    torch.save(fixture, os.path.join(GOLDEN_DIR, name + ".pt"))
    print("%-32s logits %s absmax %.4f  stages ok: %s" % (name, tuple(logits.shape), logits.abs().max(),
                                                         ",".join(hooked)))


def run_resnext_cases():
    """ResNeXt-3D (resnext3D.py; exported by pretorched/__init__.py:66-72): plain ``**kwargs`` factories and an ``fc`` head, so
    the fixtures get their own kind and their own tests (tests/test_oracle_golden.py, tests/test_gpu_resnext.py)."""
    for name, (arch, kwargs, shape) in {
            "resnext3d50_b1_t8_64": ("resnext3d50", dict(num_classes=400), (1, 3, 8, 64, 64)),
            # (cardinality != 32 is broken upstream: fc is sized cardinality * 32 * expansion, the trunk ends with 2048 channels)
            "resnext3d18_a_b2_t8_64": ("resnext3d18", dict(num_classes=10, shortcut_type='A'), (2, 3, 8, 64, 64)),
    }.items():
        RL.load()
        torch.manual_seed(SEED_INIT)
        ref = RL.build(arch, **kwargs)
        OF.randomize_bn_(ref, SEED_BN)
        ref.eval()
        x = OF.seeded_input(shape, SEED_INPUT)
        hooked, handles = {}, []
        for stage in ("maxpool", "layer1", "layer2", "layer3", "layer4"):
            handles.append(getattr(ref, stage).register_forward_hook(
                lambda m, i, o, stage=stage: hooked.__setitem__(stage, o.detach().clone())))
        with torch.no_grad():
            logits = ref(x)
        for h in handles:
            h.remove()
        hooked["logits"] = logits
        sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
        spec = dict(OF.ARCHS["resnext3d50"], layers={"resnext3d50": [3, 4, 6, 3], "resnext3d18": [2, 2, 2, 2]}[arch],
                    shortcut=kwargs.get("shortcut_type", "B"), cardinality=kwargs.get("cardinality", 32))
        stages = {}
        with torch.no_grad():
            out = OF.forward(x, sd, spec, stages)
        for k, v in hooked.items():
            assert torch.equal(stages[k], v), "oracle restatement differs from the reference at %s/%s" % (name, k)
        assert torch.equal(out, logits)
        torch.save(dict(kind="resnext", arch=arch, kwargs=kwargs, spec=spec, input_shape=tuple(shape),
                        seeds=dict(init=SEED_INIT, bn=SEED_BN, input=SEED_INPUT), logits=logits.clone(),
                        stages={k: summarize(v) for k, v in hooked.items()}, weight_digest=OF.state_digest(sd), n_state=len(sd),
                        torch_version=torch.__version__), os.path.join(GOLDEN_DIR, name + ".pt"))
        print("%-32s logits %s absmax %.4f  stages ok: %s" % (name, tuple(logits.shape), logits.abs().max(), ",".join(hooked)))


def run_relation_cases():
    trn = RL.load_trn()
    # (a) single Relation, small and at the TRN-wired size (trn.py:230-233: T=8, F=2048, bottleneck 512)
    for name, (T, Fdim, out, bott, B) in {"relation_small": (8, 256, 64, 128, 5),
                                          "relation_htrn": (8, 2048, 1024, 512, 3)}.items():
        torch.manual_seed(SEED_INIT)
        ref = trn.Relation(T, Fdim, out, bottleneck_dim=bott).eval()
        x = OF.seeded_input((B, T, Fdim), SEED_INPUT)
        with torch.no_grad():
            y = ref(x)
            y_or = OF.relation(x, ref.state_dict(), "", T, Fdim)
        assert torch.equal(y, y_or)
        torch.save(dict(kind="relation", T=T, F=Fdim, out=out, bottleneck=bott, B=B,
                        seeds=dict(init=SEED_INIT, input=SEED_INPUT), output=y.clone(),
                        weight_digest=OF.state_digest(ref.state_dict())), os.path.join(GOLDEN_DIR, name + ".pt"))
        print("%-32s out %s" % (name, tuple(y.shape)))
    # (b) MultiScaleRelation; the reference draws frame tuples from NumPy's global RNG on every forward
    T, Fdim, out, bott, B = 8, 256, 64, 128, 4
    torch.manual_seed(SEED_INIT)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):       # the constructor prints (trn.py:97-98)
        ref = trn.MultiScaleRelation(T, Fdim, out, bottleneck_dim=bott).eval()
    x = OF.seeded_input((B, T, Fdim), SEED_INPUT)
    with torch.no_grad():
        np.random.seed(123)
        y = ref(x)
        np.random.seed(123)
        y_or = OF.multiscale_relation(x, ref.state_dict(), T, Fdim)
    assert torch.equal(y, y_or)
    torch.save(dict(kind="msrelation", T=T, F=Fdim, out=out, bottleneck=bott, B=B, np_seed=123,
                    seeds=dict(init=SEED_INIT, input=SEED_INPUT), output=y.clone(),
                    weight_digest=OF.state_digest(ref.state_dict())), os.path.join(GOLDEN_DIR, "msrelation_small.pt"))
    print("%-32s out %s" % ("msrelation_small", tuple(y.shape)))


NLBLOCK_CASES = {
    # name -> (dimension, mode, sub_sample, bn_layer, channels, input shape)
    "nlblock3d_gaussian_sub": (3, "gaussian", True, True, 128, (2, 128, 4, 8, 8)),
    "nlblock3d_dot_product": (3, "dot_product", False, True, 128, (2, 128, 2, 7, 7)),
    "nlblock3d_embedded_sub": (3, "embedded_gaussian", True, True, 128, (1, 128, 4, 8, 8)),
    "nlblock2d_embedded": (2, "embedded_gaussian", False, True, 128, (2, 128, 12, 12)),
    "nlblock2d_dot_sub_nobn": (2, "dot_product", True, False, 128, (2, 128, 12, 12)),
    "nlblock1d_gaussian": (1, "gaussian", False, True, 128, (3, 128, 50)),
}


def condition_nlblock_(blk, seed):
    """Deterministic, benign parameters for a bare non-local block: BN statistics randomised, projections scaled so
    the logits stay O(1), the zero-initialised output projection (nonlocalnet.py:95-102) replaced by seeded noise."""
    OF.randomize_bn_(blk, seed)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for name, prm in blk.named_parameters():
            if name.startswith(("theta", "phi")) and name.endswith("weight"):
                prm.mul_(0.3)
            if name in ("W.weight", "W.bias"):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.05)
    return blk


def run_nlblock_cases():
    ref_nl = RL.load().models.nonlocalnet
    for name, (dim, mode, sub, bn, C, shape) in NLBLOCK_CASES.items():
        cls = getattr(ref_nl, "NonLocalBlock%dD" % dim)
        torch.manual_seed(SEED_INIT)
        ref = condition_nlblock_(cls(C, mode=mode, sub_sample=sub, bn_layer=bn), SEED_BN).eval()
        x = OF.seeded_input(shape, SEED_INPUT) * (0.3 if mode == "gaussian" else 1.0)
        with torch.no_grad():
            y = ref(x)
            y_or = OF.nonlocal_block_nd(x, ref.state_dict(), "", dim, mode, sub, bn)
        assert torch.equal(y, y_or), "oracle restatement differs from the reference for %s" % name
        torch.save(dict(kind="nlblock", dimension=dim, mode=mode, sub_sample=sub, bn_layer=bn, channels=C,
                        input_shape=tuple(shape), input_scale=(0.3 if mode == "gaussian" else 1.0),
                        seeds=dict(init=SEED_INIT, bn=SEED_BN, input=SEED_INPUT), output=y.clone(),
                        weight_digest=OF.state_digest(ref.state_dict())), os.path.join(GOLDEN_DIR, name + ".pt"))
        print("%-32s out %s absmax %.4f" % (name, tuple(y.shape), y.abs().max()))


SLOWFAST_CASES = {
    # name -> (factory, mode, layers, bottleneck, kwargs, input shape)
    "slowfast50_sf_b1_t32_64": ("resnet50", "sf", [3, 4, 6, 3], True, dict(num_classes=12), (1, 3, 32, 64, 64)),
    "slowfast18_sf_b2_t32_64": ("resnet18", "sf", [2, 2, 2, 2], False, dict(num_classes=12), (2, 3, 32, 64, 64)),
    "slowonly50_b1_t32_64": ("resnet50", "s", [3, 4, 6, 3], True, dict(num_classes=12), (1, 3, 32, 64, 64)),
    "fastonly50_b1_t16_64": ("resnet50", "f", [3, 4, 6, 3], True, dict(num_classes=12), (1, 3, 16, 64, 64)),
}


def run_slowfast_cases():
    ref_pkg = RL.load()
    for name, (factory, mode, layers, bott, kwargs, shape) in SLOWFAST_CASES.items():
        torch.manual_seed(SEED_INIT)
        ref = getattr(ref_pkg.slowfast, factory)(mode=mode, **kwargs)
        OF.randomize_bn_(ref, SEED_BN)
        ref.eval()
        x = OF.seeded_input(shape, SEED_INPUT)
        with torch.no_grad():
            y = ref(x)
            sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
            y_or = OF.slowfast_forward(x, sd, layers, bott, mode)
        assert torch.equal(y, y_or), "oracle restatement differs from the reference for %s" % name
        torch.save(dict(kind="slowfast", factory=factory, mode=mode, layers=layers, bottleneck=bott, kwargs=kwargs,
                        input_shape=tuple(shape), seeds=dict(init=SEED_INIT, bn=SEED_BN, input=SEED_INPUT),
                        logits=y.clone(), weight_digest=OF.state_digest(sd), n_state=len(sd)),
                   os.path.join(GOLDEN_DIR, name + ".pt"))
        print("%-32s logits %s absmax %.4f" % (name, tuple(y.shape), y.abs().max()))


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    # R(2+1)D first: see reference_loader.load_r2plus1d
    order = sorted(MODEL_CASES, key=lambda n: 0 if MODEL_CASES[n][0].startswith("r2plus1d") else 1)
    for name in order:
        run_model_case(name, *MODEL_CASES[name])
    run_relation_cases()
    run_slowfast_cases()
    run_nlblock_cases()
    run_resnext_cases()


if __name__ == "__main__":
    main()
