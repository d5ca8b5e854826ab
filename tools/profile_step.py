"""One steady-state G+D train step between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/profile_step.py
(the same step bench.py times: batch 64, 8 objects, 128x128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layout2img_b200.model.rcnn_discriminator_app import CombineDiscriminator128_app
from layout2img_b200.model.resnet_generator_app_v2 import ResnetGenerator128_context
from layout2img_b200.synth import make_state, schema_of, synthetic_layout
from layout2img_b200.train import make_optimizers, train_step

B = int(os.environ.get("L2I_BATCH", "64"))
warm = int(os.environ.get("L2I_WARM", "3"))
dev = torch.device("cuda:0")
G = ResnetGenerator128_context(num_classes=184, output_dim=3)
D = CombineDiscriminator128_app(num_classes=184)
G.load_state_dict(make_state(schema_of(G), 1)); D.load_state_dict(make_state(schema_of(D), 2))
G.to(dev).train(); D.to(dev).train()
# the step bench.py records into its CUDA graph, issued call by call so that the profiler sees every launch: fixed-shape
# discriminator (device-side ROI preparation, no host sync), gradients in flat buckets, capturable Adam
from layout2img_b200.train import GradBuckets
D.static_shapes = True
g_opt, d_opt = make_optimizers(G, D, capturable=True)
sg, sd = GradBuckets(G), GradBuckets(D)
d = {k: v.to(dev) for k, v in synthetic_layout(B, 8, 184, seed=0).items()}
step = lambda: train_step(G, D, g_opt, d_opt, d["real"], d["label"], d["bbox"], d["z"], d["z_im"], sync_g=sg, sync_d=sd)
for _ in range(warm):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
