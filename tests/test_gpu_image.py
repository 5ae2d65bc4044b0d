"""GPU (-m gpu): the device-side TransformImage (csrc/b2_image.cu) is bit-exact with the reference pipeline's output and with the
CPU oracle (oracle/image.py, pinned to Pillow + torchvision), and BASELINE.json configs[0] -- resnet18 on the reference's cat.jpg
-- matches the reference's logits end to end (examples/imagenet_logits.py:38-68)."""
import os

import numpy as np
import pytest
import torch

from oracle import functional as OF
from oracle import image as OI
import pretorched_x_b200 as P
from pretorched_x_b200 import transforms as TR, ops

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "resnet18_cat_224.pt")


def test_gpu_transform_is_bit_exact_and_resnet18_matches_reference_on_cat():
    fx = torch.load(GOLDEN, weights_only=False)
    dev = torch.device("cuda:0")
    tf = TR.TransformImage(fx["settings"])
    x = tf(fx["image_u8"])
    s = fx["settings"]
    want = torch.from_numpy(OI.transform_image(fx["image_u8"].numpy(), s["input_size"], s["input_space"], s["input_range"], s["mean"], s["std"]))
    assert x.is_cuda and torch.equal(x.cpu(), want)                                   # bit-exact fp32
    ref = fx["input"]
    assert torch.equal(x.cpu().reshape(-1)[::ref["step"]][:ref["sample"].numel()], ref["sample"])   # == the reference's own tensor
    m = OF.build_package_model(P, dict(fx, kind="model")).to(dev)
    with torch.no_grad():
        y = m(x.unsqueeze(0))
        a4 = tf.stem_input(fx["image_u8"])                                             # fused path: fp16 NDHWC4 straight into the stem
        y4 = m.logits(m.features_act(a4))
    scale = fx["logits"].abs().max().item()
    assert (y.cpu() - fx["logits"]).abs().max().item() <= 5e-3 * scale
    assert int(y.argmax(1)) == int(fx["logits"].argmax(1))
    assert torch.equal(y, y4)                                                          # same fp16 rounding of the same fp32 values


@pytest.mark.parametrize("h,w,kw", [(301, 123, dict()), (97, 211, dict(scale=0.7)), (256, 320, dict()),
                                    (384, 480, dict(preserve_aspect_ratio=False)), (224, 224, dict(scale=1.0))])
def test_gpu_transform_shapes_and_options(h, w, kw):
    from tests.test_image_cpu import synthetic_image
    a = synthetic_image(h, w, 5)
    for space, rng in (("RGB", [0, 1]), ("BGR", [0, 255])):
        opts = dict(input_size=[3, 224, 224], input_space=space, input_range=rng, mean=[0.4, 0.5, 0.6], std=[0.2, 0.3, 0.25])
        tf = TR.TransformImage(opts, **kw)
        got = tf(a).cpu()
        if kw.get("preserve_aspect_ratio", True):
            want = torch.from_numpy(OI.transform_image(a, opts["input_size"], space, rng, opts["mean"], opts["std"], kw.get("scale", 0.875)))
        else:
            r = OI.pil_resize_bilinear(a, 256, 256)[16:240, 16:240]
            t = r.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
            if space == "BGR":
                t = t[::-1].copy()
            if max(rng) == 255:
                t = t * np.float32(255)
            want = torch.from_numpy((t - np.asarray(opts["mean"], np.float32).reshape(3, 1, 1)) / np.asarray(opts["std"], np.float32).reshape(3, 1, 1))
        assert torch.equal(got, want), (space, (got - want).abs().max().item())


def test_gpu_random_crop_and_flips_match_torchvision():
    from PIL import Image
    import torchvision.transforms as T
    from tests.test_image_cpu import synthetic_image
    a = synthetic_image(300, 400, 9)
    opts = dict(input_size=[3, 224, 224], input_space='RGB', input_range=[0, 1], mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    ref = T.Compose([T.Resize(256), T.RandomCrop(224), T.RandomHorizontalFlip(), T.RandomVerticalFlip(), T.ToTensor(),
                     T.Normalize(mean=opts["mean"], std=opts["std"])])
    tf = TR.TransformImage(opts, random_crop=True, random_hflip=True, random_vflip=True)
    for seed in range(4):
        torch.manual_seed(seed)
        want = ref(Image.fromarray(a))
        torch.manual_seed(seed)
        got = tf(a).cpu()
        assert torch.equal(got, want), seed


def test_uint8_clip_input_matches_fp32_pipeline():
    """ClipToStemInput: uint8 THWC frames -> normalised fp16 NDHWC4 on the device == the fp32 NCDHW tensor a host pipeline would
    build (ToTensor + Normalize per frame), so both give the same logits."""
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = OF.randomize_bn_(P.resnet3d18(num_classes=11, pretrained=None), 1).eval().to(dev)
    s = P.pretrained_settings["resnet3d18"]["kinetics-400"]
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (2, 8, 64, 64, 3), generator=g, dtype=torch.uint8)
    mean = torch.tensor(s["mean"], dtype=torch.float32).view(1, 1, 1, 1, 3)
    std = torch.tensor(s["std"], dtype=torch.float32).view(1, 1, 1, 1, 3)
    x = ((u8.float() / 255.0 - mean) / std).permute(0, 4, 1, 2, 3).contiguous()          # fp32 NCDHW, computed on the host
    tf = TR.ClipToStemInput(s)
    with torch.no_grad():
        a = tf(u8.to(dev))
        want_in = ops.from_ncdhw(x.to(dev))
        assert torch.equal(a.data, want_in.data)                                          # identical fp16 NDHWC4 stem input
        assert torch.equal(m(a), m(x.to(dev)))
