"""world_size-2 gloo test of the data-parallel plumbing: after the single
all-reduce over the flat gradient bucket, both ranks hold the SUM and the optimiser's grad_scale
is 1/world.  CPU only (gloo); the collective on the GPU box is the same call over NCCL."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from segan_pytorch_b200.segan.models.model import allreduce_grads

    class Eng:
        pass
    e = Eng()
    e.grad = torch.full((1000,), float(rank + 1))
    scale = allreduce_grads(e)
    q.put((rank, float(e.grad[0]), float(e.grad.sum()), scale))
    dist.destroy_process_group()


def test_allreduce_flat_bucket_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(60)
    for rank, first, total, scale in res:
        assert first == 3.0 and total == 3000.0 and scale == 0.5


def test_rank_sharded_synthetic_data_is_disjoint():
    from segan_pytorch_b200.segan.datasets import SyntheticSEDataset
    a = SyntheticSEDataset(4, seed=111 + 0)
    b = SyntheticSEDataset(4, seed=111 + 1)
    assert not torch.equal(a.clean, b.clean)
    assert float(a.clean.abs().max()) <= 1.0
