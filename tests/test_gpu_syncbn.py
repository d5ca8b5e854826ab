"""Cross-rank batch statistics (ops.set_sync_bn -- the reference's multi-GPU SynchronizedBatchNorm2d,
model/sync_batchnorm/batchnorm.py:90-125) at operator level: a generator ResBlock (two ISLA norms, mask head with an
affine batch norm) and the PSP head (its bottleneck norm) run on TWO ranks, each with half of a batch, must reproduce
the single-process whole-batch forward, input gradients, parameter gradients (summed over ranks) and running statistics.
Operator level on purpose: no deep ReLU / 1/(sum m + 1e-6) chain amplifies rounding, so the tolerance is tight.

Two processes share cuda:0 over gloo (works on the one-GPU test box); with >= 2 GPUs the same check also runs over
NCCL, one rank per GPU.
"""
import os
import socket
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_block(psp, sl, dev, sync):
    """One ResBlock forward + backward on the batch slice `sl`; returns outputs, input grads, param grads, running stats."""
    import copy
    from layout2img_b200 import ops
    from layout2img_b200.model.resnet_generator_app_v2 import ResBlock
    torch.manual_seed(1234 + int(psp))
    B, O, CIN, COUT, H, NW = 4, 3, 64, 32, 8, 40
    blk = ResBlock(CIN, COUT, upsample=True, num_w=NW, psp_module=psp)
    with torch.no_grad():
        for n, p in blk.named_parameters():
            if n.endswith("bias") or p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    blk = copy.deepcopy(blk).to(dev).train()
    if psp:
        for st in blk.conv_mask[0].stages:
            st[2].eval()      # plain nn.BatchNorm2d: the reference does not synchronise these either
        blk.conv_mask[0].dropout_mask = torch.ones(sl.stop - sl.start, 100)
    g = torch.Generator().manual_seed(99)
    x = torch.randn(B, H, H, CIN, generator=g)
    w = torch.randn(B * O, NW, generator=g)
    bb = torch.rand(B, O, 16, 16, generator=g)
    p1 = torch.randn(B, 2 * H, 2 * H, COUT, generator=g)
    p2 = torch.randn(B, 2 * H, 2 * H, 184, generator=g)
    xs = x[sl].to(dev).requires_grad_()
    ws = w[sl.start * O:sl.stop * O].to(dev).requires_grad_()
    bs = bb[sl].to(dev).requires_grad_()
    ops.set_sync_bn(sync)
    try:
        out, mask = blk(xs, ws, bs)
        ((out * p1[sl].to(dev)).sum() + (mask * p2[sl].to(dev)).sum()).backward()
    finally:
        ops.set_sync_bn(False)
    res = {"out": out.detach(), "mask": mask.detach(), "dx": xs.grad, "dw_vec": ws.grad, "dbbox": bs.grad}
    pg = {n: p.grad.detach().clone() for n, p in blk.named_parameters() if p.grad is not None}
    rs = {n: v.detach().clone() for n, v in blk.state_dict().items() if "running_" in n and "stages" not in n}
    return res, pg, rs


def _worker(rank, world, port, backend, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dev = torch.device("cuda", rank if backend == "nccl" else 0)
        torch.cuda.set_device(dev)
        if backend == "nccl":
            dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        else:
            dist.init_process_group("gloo", rank=rank, world_size=world)
        msgs = []
        for psp in (False, True):
            half = 2
            sl = slice(rank * half, (rank + 1) * half)
            res, pg, rs = _run_block(psp, sl, dev, True)
            for t in pg.values():                              # data-parallel gradient = sum over the ranks' shards
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ref, rpg, rrs = _run_block(psp, slice(0, 2 * half), dev, False)

            def rel(a, b):
                return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-20)).item()

            for k, v in res.items():
                want = ref[k][sl.start * 3:sl.stop * 3] if k == "dw_vec" else ref[k][sl]      # the latents are (b * O, num_w)
                e = rel(v, want)
                if e > 2e-4:
                    msgs.append(f"psp={psp} {k}: {e:.2e}")
            for k, v in pg.items():
                if k.endswith(("conv1.bias", "conv_mask.0.bias")) or rpg[k].abs().max().item() < 1e-6:
                    continue                                   # conv biases in front of a batch norm: the true gradient is 0
                e = rel(v, rpg[k])
                if e > 5e-4:
                    msgs.append(f"psp={psp} grad {k}: {e:.2e}")
            for k, v in rs.items():
                e = (v - rrs[k]).abs().max().item()
                if e > 1e-5:
                    msgs.append(f"psp={psp} {k}: {e:.2e}")
            if len(rs) < 3:
                msgs.append(f"psp={psp}: only {len(rs)} running-stat buffers compared")
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, msgs))
    except Exception:
        q.put((rank, ["EXC " + traceback.format_exc()]))


def _launch(backend):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, backend, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    for r in range(world):
        assert not out[r], f"rank {r} ({backend}): " + "; ".join(out[r])


def test_sync_bn_two_ranks_one_gpu_gloo():
    _launch("gloo")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sync_bn_two_ranks_nccl():
    _launch("nccl")
