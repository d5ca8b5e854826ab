"""Run ONE training iteration of the UNMODIFIED reference modules (baseline/_ref, staged by oracle/make_ref.py) on
cuda:0 in fp32 with TF32 off, on the seeded inputs / weights the parity tests use, and save every result.

    python tests/ref_gpu_runner.py OUT.pt BATCH NUM_OBJ NUM_CLASSES DATA_SEED N_PAD SEED_G SEED_D DROPOUT_SEED

Separate process: the reference's top-level package is called `model`, like this repository's drop-in shim.
Test infrastructure only.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    out_path = sys.argv[1]
    batch, num_obj, ncls, seed, n_pad, seed_g, seed_d, dseed = (int(a) for a in sys.argv[2:10])
    from oracle import ref_harness as R
    Gc, Dc = R.import_reference(cpu=False)
    from layout2img_b200.synth import make_state, schema_of, synthetic_layout
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    G, D = Gc(num_classes=ncls, output_dim=3), Dc(num_classes=ncls)
    G.load_state_dict(make_state(schema_of(G), seed_g)); D.load_state_dict(make_state(schema_of(D), seed_d))
    G.to(dev).train(); D.to(dev).train()
    data = synthetic_layout(batch, num_obj, ncls, seed=seed, n_pad=n_pad)
    g_opt, d_opt = R.make_optimizers(G, D)
    res, cap = {}, {}

    def drop_hook(mod, inp, out):         # the keep-mask the reference's Dropout2d drew (channels that are all zero: 0)
        x, y = inp[0].detach(), out.detach()
        ratio = torch.where(x > 0, y / x.clamp_min(1e-30), torch.zeros_like(x))
        cap["keep"] = ratio.amax(dim=(2, 3)).cpu()

    h = G.res4.conv_mask[0].bottleneck[3].register_forward_hook(drop_hook)

    def record(tag):
        net = D if tag == "d" else G
        for n, p in net.named_parameters():
            res[tag + "." + n] = p.grad.detach().clone().cpu()

    torch.manual_seed(dseed)
    dl, gl, fake = R.train_step(G, D, g_opt, d_opt, data["real"].to(dev), data["label"].to(dev), data["bbox"],
                                data["z"].to(dev), data["z_im"].to(dev), record=record)
    h.remove()
    res.update(d_loss=dl.cpu(), g_loss=gl.cpu(), fake=fake.cpu(), keep=cap["keep"],
               PG={k: v.detach().cpu() for k, v in G.state_dict().items()},
               PD={k: v.detach().cpu() for k, v in D.state_dict().items()})
    torch.save(res, out_path)
    print("reference step done:", float(dl), float(gl))


if __name__ == "__main__":
    main()
