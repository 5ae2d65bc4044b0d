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


def _gan_worker(rank, world, port, ret):
    """The generator workload shards z / class ids like clips and has NO data-path collective: after the one broadcast of
    parameters and buffers (spectral-norm vectors, standing statistics, the 0-dim attention gate) every rank holds the same
    generator, and the per-rank shards tile the batch exactly."""
    import pretorched_x_b200 as P
    from oracle import biggan as OB
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_from_env(backend="gloo")
    torch.manual_seed(200 + rank)                       # ranks start with different weights and buffers
    G = P.biggan_deep(128, G_ch=16, n_classes=10, G_init="N02")
    with torch.no_grad():
        G.blocks[3][2].gamma.fill_(0.25 + rank)
        G.blocks[0][0].bn1.stored_var.mul_(1.0 + rank)
    parallel.broadcast_parameters(G)
    torch.manual_seed(200)
    ref = P.biggan_deep(128, G_ch=16, n_classes=10, G_init="N02")
    with torch.no_grad():
        ref.blocks[3][2].gamma.fill_(0.25)
    ok = all(torch.equal(a, b) for a, b in zip(G.state_dict().values(), ref.state_dict().values()))
    z, labels = OB.seeded_inputs(5, 10, 3)
    zs, ls = parallel.shard_batch(z), parallel.shard_batch(labels)
    lo, hi = parallel.shard_bounds(5, world, rank)
    ok = ok and torch.equal(zs, z[lo:hi]) and torch.equal(ls, labels[lo:hi])
    # the restatement on the shard equals the shard of the restatement on the full batch (samples are independent)
    sd = G.state_dict()
    with torch.no_grad():
        full = OB.generator_forward(z, labels, sd, 128, 16)
        mine = OB.generator_forward(zs, ls, sd, 128, 16)
    ok = ok and torch.allclose(mine, full[lo:hi], atol=1e-5)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_generator_replication_and_sharding():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gan_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    assert ret[0] and ret[1], dict(ret)
