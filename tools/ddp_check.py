"""Numerical check of the data-parallel path on 2 GPUs (run under torchrun --nproc-per-node 2):
the discriminator has no batch-coupled layer, so the all-reduced (mean) gradient of two ranks that each
take half of a batch must equal the gradient of one process on the whole batch (equal object counts per shard).
Prints max relative deviation per parameter group and exits non-zero on mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from layout2img_b200.model.rcnn_discriminator_app import CombineDiscriminator128_app
from layout2img_b200.synth import make_state, schema_of, synthetic_layout
from layout2img_b200.train import GradAllReducer, d_loss_fn

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
    loss = d_loss_fn(D(d["real"], d["bbox"], d["label"].unsqueeze(-1)), D(fake[sl].to(dev), d["bbox"], d["label"].unsqueeze(-1)))
    loss.backward()
    if sync:
        GradAllReducer(D)()
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
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if worst < 2e-2 else 1)
