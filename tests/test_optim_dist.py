"""Host-side logic of the data-parallel path on CPU: zero-copy gradient buckets with hook-driven all-reduces, the
rank-0 broadcast and the global-object-count loss normalisation over a world_size-2 gloo group (the NCCL path on the
B200 box uses the same code with a side stream)."""
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
    from layout2img_b200.train import GradBuckets, _obj_mean, global_object_scale
    # ---- 1. bucketed all-reduce driven by the post-accumulate hooks; replicas start DIFFERENT and are broadcast from rank 0
    torch.manual_seed(rank)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    net.add_module("unused", torch.nn.Linear(2, 2))            # never receives a gradient: its bucket is reduced as zeros
    buckets = GradBuckets(net, bucket_mb=1e-4)                  # tiny buckets: several all-reduces per backward
    params0 = [p.detach().clone() for p in net.parameters()]
    x = torch.arange(10, dtype=torch.float32).view(2, 5) / 10.0
    layouts = []
    for _ in range(3):                                          # views stay attached, counters reset; the second zero_grad
        buckets.zero_grad()                                     # re-cuts the buckets in the completion order seen in the first pass
        layouts.append((list(buckets.order), [tuple(b["members"]) for b in buckets.buckets], buckets.reordered))
        ((rank + 1.0) * net[1](net[0](x))).sum().backward()
        buckets.finish()
    grads = [p.grad.clone() for p in net.parameters()]
    is_view = all(p.grad.untyped_storage().data_ptr() == buckets.flat.untyped_storage().data_ptr() for p in net.parameters())
    # ---- 1b. the one-call set-up of a rank: buckets for both networks + global-batch BN statistics switched on
    from layout2img_b200 import ops
    from layout2img_b200.train import setup_data_parallel
    gA, gB = torch.nn.Linear(3, 3), torch.nn.Linear(3, 2)
    sA, sB = setup_data_parallel(gA, gB)
    dp_ok = (sA.world == sB.world == world and ops._SYNC_BN["world"] == world and sA.params[0] is gA.weight
             and sB.params[0] is gB.weight)
    ops.set_sync_bn(False)
    dp_ok = dp_ok and ops._SYNC_BN["world"] == 1
    # ---- 2. object-loss normalisation over the GLOBAL object count (ranks hold different numbers of valid objects)
    k_r = 3 if rank == 0 else 5
    label = torch.cat([torch.ones(k_r, dtype=torch.long), torch.zeros(8 - k_r, dtype=torch.long)]).view(1, 8)
    scale = global_object_scale(label)
    w = torch.nn.Parameter(torch.tensor([0.5, -0.25]))
    feats = torch.arange(16, dtype=torch.float32).view(8, 2)[rank * 3: rank * 3 + k_r] / 7.0
    _obj_mean(torch.relu(1.0 - feats @ w), scale).backward()
    gw = w.grad.clone()
    dist.all_reduce(gw)
    gw /= world                                                 # what the gradient average over the ranks yields
    # by value (numpy), not as shared-memory handles: a torch tensor in an mp queue must outlive its receiver's unpickling,
    # and this process exits right after the barrier
    q.put((rank, [t.numpy() for t in params0], [t.numpy() for t in grads], is_view and dp_ok, gw.numpy(), layouts))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_buckets_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(world):
        r, params0, grads, is_view, gw, layouts = q.get(timeout=120)
        res[r] = [[torch.from_numpy(a) for a in params0], [torch.from_numpy(a) for a in grads], is_view, torch.from_numpy(gw), layouts]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # replicas were broadcast from rank 0
    for a, b in zip(res[0][0], res[1][0]):
        assert torch.equal(a, b)
    # expected: mean over ranks of (rank + 1) * g = 1.5 * g, g = the single-process gradient on rank 0's weights
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    x = torch.arange(10, dtype=torch.float32).view(2, 5) / 10.0
    net(x).sum().backward()
    want = [1.5 * p.grad for p in net.parameters()] + [torch.zeros(2, 2), torch.zeros(2)]
    for r in range(world):
        assert res[r][2], "gradients must be views into the flat buffer"
        for got, w in zip(res[r][1], want):
            assert torch.allclose(got, w, rtol=1e-5, atol=1e-6), (r, got, w)
    # bucket layout: reverse registration order at first; from the second iteration on the observed completion order
    # (net[1] before net[0], bias before weight as autograd finishes them; the unused layer keeps its place at the end),
    # identical on both ranks, the tail bucket holding the last-finished parameters
    for r in range(world):
        first, second, third = res[r][4]
        assert first[0] == [5, 4, 3, 2, 1, 0] and first[2] is False
        assert second[2] is True and second == third == res[0][4][1]
        assert set(second[0][:2]) == {2, 3} and set(second[0][2:4]) == {0, 1} and second[0][4:] == [5, 4]
        assert sorted(m for b in second[1] for m in b) == [0, 1, 2, 3, 4, 5]
    # K-weighted object loss: equals the gradient of ONE mean over all 8 valid objects of both ranks
    w = torch.nn.Parameter(torch.tensor([0.5, -0.25]))
    allf = torch.arange(16, dtype=torch.float32).view(8, 2) / 7.0
    feats = torch.cat([allf[0:3], allf[3:8]])
    torch.relu(1.0 - feats @ w).mean().backward()
    for r in range(world):
        assert torch.allclose(res[r][3], w.grad, rtol=1e-5, atol=1e-7), (res[r][3], w.grad)


def test_shards_are_disjoint_and_deterministic():
    """Rank r draws its batch shard from seed r: disjoint across ranks, reproducible."""
    from layout2img_b200.synth import synthetic_layout
    a, b, a2 = synthetic_layout(4, 8, seed=0), synthetic_layout(4, 8, seed=1), synthetic_layout(4, 8, seed=0)
    assert torch.equal(a["z"], a2["z"]) and torch.equal(a["bbox"], a2["bbox"])
    assert not torch.equal(a["z"], b["z"])
    assert a["bbox"].shape == (4, 8, 4) and a["label"].min() >= 1


def test_bucket_layout_follows_the_given_order():
    """GradBuckets._layout: the flat buffer follows the completion order, every bucket is one contiguous range of it,
    buckets close at the size limit, and the last-finished parameters (<= tail_mb) form their own final bucket."""
    from layout2img_b200.train import GradBuckets
    net = torch.nn.Sequential(*[torch.nn.Linear(16, 16, bias=False) for _ in range(6)])     # 6 x 1 KiB gradients
    b = GradBuckets(net, bucket_mb=2.5 / 1024, tail_mb=1.5 / 1024)                           # 2.5 KiB buckets, 1.5 KiB tail
    assert b.order == [5, 4, 3, 2, 1, 0] and b.reordered                                      # single process: nothing to learn
    assert [bk["members"] for bk in b.buckets] == [[5, 4, 3], [2, 1], [0]]
    order = [2, 0, 5, 1, 4, 3]
    b._layout(order)
    assert [bk["members"] for bk in b.buckets] == [[2, 0, 5], [1, 4], [3]]
    pos = 0
    for bk in b.buckets:
        s, e = bk["range"]
        assert s == pos and e - s == 256 * len(bk["members"])
        for k, m in enumerate(bk["members"]):                     # member k of a bucket sits at its k-th slice
            assert b.views[m].data_ptr() == b.flat.data_ptr() + 4 * (s + 256 * k)
        pos = e
    assert pos == b.flat.numel() == 6 * 256
    for p, v in zip(net.parameters(), b.views):
        assert p.grad is v
    net(torch.ones(1, 16)).sum().backward()                       # autograd accumulates into the views in place
    assert all(p.grad is v and float(v.abs().sum()) > 0 for p, v in zip(net.parameters(), b.views))
    b.zero_grad()
    assert float(b.flat.abs().sum()) == 0.0 and all(p.grad is v for p, v in zip(net.parameters(), b.views))
