"""Step-by-step run of one standalone non-local block with a sync after every launch (debug aid)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pretorched_x_b200 import ops, engine
from pretorched_x_b200.models.nonlocalnet import NonLocalBlock3D
from oracle import functional as OF

dev = torch.device("cuda:0")
torch.manual_seed(0)
blk = NonLocalBlock3D(256)
with torch.no_grad():
    blk.theta.weight.mul_(0.05); blk.phi.weight.mul_(0.05); blk.W[1].weight.fill_(1.0)
OF.randomize_bn_(blk, 3)
blk.eval()
x = OF.seeded_input((2, 256, 2, 7, 7), 4).half().float()
want = OF.nonlocal_block(x, {"nl." + k: v for k, v in blk.state_dict().items()}, "nl")
blk = blk.to(dev)

orig = ops._lib.check
def checked(rc, what):
    orig(rc, what)
    torch.cuda.synchronize()
    print("ok:", what, flush=True)
ops._lib.check = checked
a = ops.from_ncdhw(x.to(dev), pitch=256)
out = engine.run_nonlocal(blk, a)
got = ops.to_ncdhw(out).cpu()
print("rel err", ((got - want).abs().max() / want.abs().max()).item())
