"""Host-side logic of the data-parallel path on CPU: the flat-bucket gradient all-reduce over a
world_size-2 gloo group (the NCCL path on the B200 box uses the same code)."""
import os
import socket

import pytest
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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from layout2img_b200.train import GradAllReducer
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    # rank-dependent gradients; one parameter left without a gradient on rank 1
    for i, p in enumerate(net.parameters()):
        if rank == 1 and i == 3:
            continue
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    GradAllReducer(net)()
    out = [p.grad.clone() for p in net.parameters()]
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_allreduce_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i in range(4):
        want = (1 * (i + 1) + (0 if i == 3 else 2 * (i + 1))) / 2.0     # mean over ranks; missing grad counts as 0
        for r in range(world):
            assert torch.allclose(res[r][i], torch.full_like(res[r][i], want)), (r, i)


def test_shards_are_disjoint_and_deterministic():
    """Rank r draws its batch shard from seed r: disjoint across ranks, reproducible."""
    from layout2img_b200.synth import synthetic_layout
    a, b, a2 = synthetic_layout(4, 8, seed=0), synthetic_layout(4, 8, seed=1), synthetic_layout(4, 8, seed=0)
    assert torch.equal(a["z"], a2["z"]) and torch.equal(a["bbox"], a2["bbox"])
    assert not torch.equal(a["z"], b["z"])
    assert a["bbox"].shape == (4, 8, 4) and a["label"].min() >= 1
