"""Generate tests/golden/biggan_*.pt from oracle/biggan.py -- OUR restatement of the published BigGAN-deep generator.

    python -m oracle.make_golden_biggan

Unlike every other fixture in tests/golden (produced by the unmodified reference), these pin nothing but the
restatement against itself over time: /root/reference has no GAN code, so BigGAN
parity is "unpinned" by construction and DESIGN.md says so.  The fixtures make the oracle's behaviour a committed,
reviewable artefact and let the CPU suite notice accidental changes to it.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import biggan as OB                 # noqa: E402
from oracle import functional as OF             # noqa: E402
from oracle.make_golden import summarize        # noqa: E402
import pretorched_x_b200 as P                   # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# name -> (resolution, ch, n_classes, batch)
CASES = {
    "biggan_deep128_ch16_b4": (128, 16, 10, 4),
    "biggan_deep256_ch16_b2": (256, 16, 10, 2),
}


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name, (res, ch, ncls, B) in CASES.items():
        model, sd, z, labels = OB.build_case(P.biggan_deep, res, ch, ncls, B)
        stages = {}
        with torch.no_grad():
            img = OB.generator_forward(z, labels, sd, res, ch, stages=stages)
        torch.save(dict(kind="biggan", resolution=res, ch=ch, n_classes=ncls, batch=B, init="N02",
                        seeds=dict(init=0, input=1), image=summarize(img), stages={k: summarize(v) for k, v in stages.items()},
                        weight_digest=OF.state_digest(sd), n_state=len(sd), keys=list(sd), torch_version=torch.__version__),
                   os.path.join(GOLDEN_DIR, name + ".pt"))
        print("%-28s image %s std %.4f stages %s" % (name, tuple(img.shape), img.std(), ",".join(stages)))


if __name__ == "__main__":
    main()
