"""CPU: pin oracle/image.py (the restatement of TransformImage, pretorched/transforms/utils.py:34-81) against Pillow and
torchvision themselves and against the reference-generated fixture; check the product's host-side logic (coefficient tables,
output-size / crop rules, RNG consumption) against both.  No GPU needed."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import image as OI
from oracle import functional as OF
from pretorched_x_b200 import transforms as TR

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "resnet18_cat_224.pt")


def synthetic_image(h, w, seed):
    g = np.random.RandomState(seed)
    base = g.randint(0, 256, size=(h // 7 + 2, w // 7 + 2, 3)).astype(np.uint8)
    img = np.kron(base, np.ones((7, 7, 1), dtype=np.uint8))[:h, :w]
    return np.ascontiguousarray(img + g.randint(0, 8, size=img.shape).astype(np.uint8) // 2)


@pytest.mark.parametrize("h,w,size", [(384, 480, 256), (301, 123, 256), (97, 211, 333), (64, 64, 256), (500, 500, 64)])
def test_restated_resize_equals_pillow(h, w, size):
    from PIL import Image
    import torchvision.transforms as T
    a = synthetic_image(h, w, h + w)
    want = np.asarray(T.Resize(size)(Image.fromarray(a)))
    nh, nw = OI.resized_size(h, w, size)
    assert (nh, nw) == want.shape[:2] == TR.resized_size(h, w, size)
    assert np.array_equal(OI.pil_resize_bilinear(a, nh, nw), want)


def test_restated_pipeline_equals_torchvision_compose():
    from PIL import Image
    import torchvision.transforms as T
    a = synthetic_image(384, 480, 3)
    tf = T.Compose([T.Resize(256), T.CenterCrop(224), T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    assert torch.equal(tf(Image.fromarray(a)), torch.from_numpy(OI.transform_image(a)))


def test_restatement_matches_reference_fixture():
    """The fixture's tensor was produced by the reference's own TransformImage on data/cat.jpg (oracle/make_golden.py)."""
    fx = torch.load(GOLDEN, weights_only=False)
    s = fx["settings"]
    x = torch.from_numpy(OI.transform_image(fx["image_u8"].numpy(), s["input_size"], s["input_space"], s["input_range"], s["mean"], s["std"]))
    ref = fx["input"]
    assert tuple(x.shape) == ref["shape"]
    assert torch.equal(x.reshape(-1)[::ref["step"]][:ref["sample"].numel()], ref["sample"])
    assert OF.digests_match(OF.state_digest({"x": x}), {"x": fx["input_sha"]})
    # and the network on it (config #1: resnet18, single image)
    import pretorched_x_b200 as P
    m = OF.build_package_model(P, dict(fx, kind="model"))
    with torch.no_grad():
        out = OF.forward(x.unsqueeze(0), m.state_dict(), "resnet18")
    assert (out - fx["logits"]).abs().max().item() <= 1e-5 * fx["logits"].abs().max().item()


@pytest.mark.parametrize("a,b", [(480, 320), (384, 256), (123, 333), (301, 64), (50, 200), (1000, 7)])
def test_product_coefficient_tables_equal_oracle(a, b):
    bo, kk = OI.bilinear_coeffs(a, b)
    tb, tk, ks = TR.resample_coeffs(a, b, "cpu")
    assert ks == kk.shape[1] and np.array_equal(tb.numpy(), bo) and np.array_equal(tk.numpy(), kk)


def test_plan_follows_torchvision_rules_and_rng_order():
    import torchvision.transforms as T
    opts = dict(input_size=[3, 224, 224], input_space='RGB', input_range=[0, 1], mean=[0.5] * 3, std=[0.5] * 3)
    tf = TR.TransformImage(opts, device="cpu")
    assert tf.plan(384, 480) == (256, 320, 16, 48, 224, False, False)
    assert tf.plan(301, 123)[:2] == OI.resized_size(301, 123, 256)
    tfr = TR.TransformImage(opts, random_crop=True, random_hflip=True, random_vflip=True, device="cpu")
    torch.manual_seed(11)
    rh, rw, top, left, crop, hf, vf = tfr.plan(384, 480)
    torch.manual_seed(11)
    i, j, th, tw = T.RandomCrop.get_params(torch.empty(3, rh, rw), (224, 224))
    assert (top, left) == (i, j)
    assert hf == bool(torch.rand(1) < 0.5) and vf == bool(torch.rand(1) < 0.5)
    tfn = TR.TransformImage(opts, preserve_aspect_ratio=False, device="cpu")
    assert tfn.plan(384, 480)[:2] == (int(224 / 0.875), int(224 / 0.875))
