"""CPU: the N>1 data-parallel path (shard -> replica forward -> all-gather of logits) on world_size-2 gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pretorched_x_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_from_env(backend="gloo")
    torch.manual_seed(100 + rank)                       # ranks start with different weights
    model = torch.nn.Linear(6, 5)
    dp = parallel.DataParallelForward(model)            # broadcast from rank 0
    full = torch.arange(total * 6, dtype=torch.float32).view(total, 6) / 10.0
    with torch.no_grad():
        out = dp(full)
    torch.manual_seed(100)
    ref = torch.nn.Linear(6, 5)
    with torch.no_grad():
        want = ref(full)
    ok = out.shape == want.shape and torch.allclose(out, want, atol=1e-6)
    lo, hi = parallel.shard_bounds(total, world, rank)
    ok = ok and parallel.shard_batch(full).shape[0] == hi - lo
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    for total in (4, 5):                                # even and ragged split
        mgr = mp.Manager()
        ret = mgr.dict()
        port = _free_port()
        mp.spawn(_worker, args=(2, port, total, ret), nprocs=2, join=True)
        assert ret[0] and ret[1], (total, dict(ret))
