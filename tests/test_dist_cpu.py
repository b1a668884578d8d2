"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sharding the minibatch across ranks and averaging
the summed gradients reproduces the single-process gradients on the concatenated batch, and the
communicator bootstrap (rank-0 id -> broadcast) delivers identical bytes to every rank."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import cyclegan_oracle as O
    # communicator bootstrap exactly as CycleGAN._attach_communicator does it
    obj = [bytes(range(128)) if rank == 0 else b""]
    dist.broadcast_object_list(obj, src=0)
    assert obj[0] == bytes(range(128))
    P = O.init_params(seed=5, dtype=torch.float64, perturb_affine=True)
    keep = [k for k in P if k.startswith("discriminator_A/")]
    A, B = O.synthetic_batch(seed=11, batch=2 * world, frames=32, dtype=torch.float64)
    # a discriminator-only objective keeps this CPU test in seconds; the all-reduce logic is identical
    def grads(Ab):
        Pg = {k: (v.detach().clone().requires_grad_(True) if k in keep else v) for k, v in P.items()}
        d = O.discriminator_forward(Ab, Pg, "discriminator_A")
        loss = O.l2_loss(torch.ones_like(d), d)
        g = torch.autograd.grad(loss, [Pg[k] for k in keep])
        return torch.cat([x.reshape(-1) for x in g])
    local = grads(A[rank * 2:(rank + 1) * 2])
    flat = local.clone()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= world                                   # folded into Adam's grad_scale in the engine
    if rank == 0:
        full = grads(A)
        q.put(float((flat - full).norm() / full.norm()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gradient_allreduce_matches_full_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(500)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-12
