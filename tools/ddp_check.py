"""Numerical check of the data-parallel path on 2 GPUs (run under torchrun --nproc-per-node 2):
 1. the discriminator has no batch-coupled layer, so the all-reduced (mean) gradient of two ranks that each
    take half of a batch must equal the gradient of one process on the whole batch;
 2. with ops.set_sync_bn(True) (cross-rank batch statistics, the reference's multi-GPU SynchronizedBatchNorm2d)
    the generator on two half-batches must reproduce the whole-batch forward, running statistics and -- up to the
    chaotic amplification measured by a 1e-6-perturbed whole-batch run -- gradients.
Exits non-zero on mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from layout2img_b200.model.rcnn_discriminator_app import CombineDiscriminator128_app
from layout2img_b200.synth import make_state, schema_of, synthetic_layout
from layout2img_b200.train import GradBuckets, d_loss_fn

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
B = 8
data = synthetic_layout(B * world, 8, 184, seed=0)
fake = torch.rand(B * world, 3, 128, 128, generator=torch.Generator().manual_seed(1)) * 2 - 1

def grads(sl, sync):
    D = CombineDiscriminator128_app(num_classes=184)
    D.load_state_dict(make_state(schema_of(D), 2))
    D.to(dev).train()
    d = {k: v[sl].to(dev) for k, v in data.items()}
    buckets = GradBuckets(D, broadcast=False) if sync else None     # .grad = views into the flat bucket buffer
    loss = d_loss_fn(D(d["real"], d["bbox"], d["label"].unsqueeze(-1)), D(fake[sl].to(dev), d["bbox"], d["label"].unsqueeze(-1)))
    loss.backward()
    if sync:
        buckets.finish()
    return {n: p.grad.detach().clone() for n, p in D.named_parameters()}, D

g_dp, _ = grads(slice(rank * B, (rank + 1) * B), True)          # each rank: its shard, then all-reduce (mean)
g_ref, _ = grads(slice(0, B * world), False)                    # every rank: the whole batch, no collective
worst = 0.0
for n in g_ref:
    m = g_ref[n].abs().max().item()
    e = (g_dp[n] - g_ref[n]).abs().max().item() / max(m, 1e-20)
    worst = max(worst, e)
    if e > 5e-3 and rank == 0:
        print(f"  {n}: max|diff|/max|ref| = {e:.3e}")
if rank == 0:
    print(f"ddp_check: world {world}, {len(g_ref)} gradient tensors, worst max|diff|/max|ref| = {worst:.3e}")
# ---- generator with cross-rank batch statistics (ops.set_sync_bn): two half-batches == one whole batch
from layout2img_b200 import ops
from layout2img_b200.model.resnet_generator_app_v2 import ResnetGenerator128_context
proj = torch.randn(B * world, 3, 128, 128, generator=torch.Generator().manual_seed(2))

def g_run(sl, sync, perturb=0.0):
    ops.set_sync_bn(sync)
    G = ResnetGenerator128_context(num_classes=184, output_dim=3)
    G.load_state_dict(make_state(schema_of(G), 1))
    G.to(dev).train()
    for st in G.res4.conv_mask[0].stages:
        st[2].eval()      # the PSP stages use plain nn.BatchNorm2d, which the reference does not synchronise either
    n = sl.stop - sl.start
    G.res4.conv_mask[0].dropout_mask = torch.ones(n, 100)
    d = {k: v[sl].to(dev) for k, v in data.items()}
    buckets = GradBuckets(G, broadcast=False) if sync else None
    fake = G(d["z"], d["bbox"], d["z_im"] * (1.0 + perturb), d["label"])
    (fake * proj[sl].to(dev)).mean().backward()
    if sync:
        buckets.finish()
    ops.set_sync_bn(False)
    return fake.detach(), {k: p.grad.detach().clone() for k, p in G.named_parameters() if p.grad is not None}, \
        {k: v.detach().clone() for k, v in G.state_dict().items() if "running_" in k}

sl = slice(rank * B, (rank + 1) * B)
f_dp, g_dp, rs_dp = g_run(sl, True)
f_ref, g_ref, rs_ref = g_run(slice(0, B * world), False)
e_f = (f_dp - f_ref[sl]).abs().max().item()
worst_g, worst_name, rows = 0.0, "", []
for n in g_ref:
    m = g_ref[n].abs().max().item()
    if m < 1e-7:          # conv biases in front of a batch norm: the true gradient is exactly zero
        continue
    e = (g_dp[n] - g_ref[n]).abs().max().item() / m
    l2 = (g_dp[n] - g_ref[n]).norm().item() / max(g_ref[n].norm().item(), 1e-30)
    rows.append((e, m, n, l2))
    if e > worst_g:
        worst_g, worst_name = e, n
# noise baseline: the same whole-batch run with z_im perturbed by 1e-6 relative (chaotic amplification through
# ReLU kinks and ISLA's 1/(sum m + 1e-6))
_, g_pert, _ = g_run(slice(0, B * world), False, perturb=1e-6)
if rank == 0:
    for e, m, n, l2 in sorted(rows, reverse=True)[:6]:
        ep = (g_pert[n] - g_ref[n]).abs().max().item() / m
        l2p = (g_pert[n] - g_ref[n]).norm().item() / max(g_ref[n].norm().item(), 1e-30)
        print(f"  {n}: max-diff {e:.3e} (rel-L2 {l2:.3e}) of max {m:.3e} | 1e-6-perturbed whole batch: {ep:.3e} (rel-L2 {l2p:.3e})")
    worst_l2 = max(r[3] for r in rows)
    print(f"  worst rel-L2 over all gradient tensors: {worst_l2:.3e}")
# verdict on the gradients: every tensor's deviation must be explained by the chaotic-noise baseline
ratio_bad = [n for e, m, n, l2 in rows
             if l2 > 3.0 * (g_pert[n] - g_ref[n]).norm().item() / max(g_ref[n].norm().item(), 1e-30) + 2e-3]
if rank == 0 and ratio_bad:
    print("  tensors beyond 3x the perturbation baseline:", ratio_bad[:8])
e_rs = max((rs_dp[k] - rs_ref[k]).abs().max().item() for k in rs_ref if k.startswith(("res", "final")) and "stages" not in k)
if rank == 0:
    print(f"ddp_check sync-BN generator: max|fake diff| = {e_f:.3e}; worst gradient tensor {worst_name}: {worst_g:.3e} of its max; "
          f"max running-stat diff {e_rs:.3e}")
ok = worst < 4e-2 and e_f < 5e-4 and not ratio_bad and e_rs < 1e-4
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
