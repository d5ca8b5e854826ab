"""Secondary BASELINE.json configurations (bench.py measures the headline one, configs[2]/[3]):
   configs[1]  generator-only forward, batch 64, 128x128, 8 objects   (eval mode, no_grad)
   configs[4]  VG-shape G+D train step, batch 32, 16 objects, 179 classes
   python tools/bench_configs.py [--steps K] [--warmup W]      -> one JSON line per configuration"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layout2img_b200.model.rcnn_discriminator_app import CombineDiscriminator128_app
from layout2img_b200.model.resnet_generator_app_v2 import ResnetGenerator128_context
from layout2img_b200.synth import make_state, schema_of, synthetic_layout
from layout2img_b200.train import make_optimizers, train_step


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    # configs[1]
    G = ResnetGenerator128_context(num_classes=184, output_dim=3)
    G.load_state_dict(make_state(schema_of(G), 1)); G.to(dev).eval()
    d = {k: v.to(dev) for k, v in synthetic_layout(64, 8, 184, seed=0).items()}
    with torch.no_grad():
        ms = timed(lambda: G(d["z"], d["bbox"], d["z_im"], d["label"]), a.steps, a.warmup)
    print(json.dumps({"metric": "generator-only forward images/sec @128x128, batch 64, 8 objects (BASELINE configs[1])",
                      "value": 64 / (ms / 1e3), "unit": "images/sec", "ms_per_step": ms, "steps": a.steps, "warmup": a.warmup,
                      "dtype": "fp32", "data": "synthetic"}), flush=True)
    # configs[4]
    G = ResnetGenerator128_context(num_classes=179, output_dim=3)
    D = CombineDiscriminator128_app(num_classes=179)
    G.load_state_dict(make_state(schema_of(G), 1)); D.load_state_dict(make_state(schema_of(D), 2))
    G.to(dev).train(); D.to(dev).train()
    g_opt, d_opt = make_optimizers(G, D)
    d = {k: v.to(dev) for k, v in synthetic_layout(32, 16, 179, seed=0).items()}
    ms = timed(lambda: train_step(G, D, g_opt, d_opt, d["real"], d["label"], d["bbox"], d["z"], d["z_im"]), a.steps, a.warmup)
    print(json.dumps({"metric": "VG-shape G+D train-step images/sec @128x128, batch 32, 16 objects, 179 classes (BASELINE configs[4])",
                      "value": 32 / (ms / 1e3), "unit": "images/sec", "ms_per_step": ms, "steps": a.steps, "warmup": a.warmup,
                      "dtype": "fp32", "data": "synthetic"}), flush=True)


if __name__ == "__main__":
    main()
