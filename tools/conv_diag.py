"""GPU diagnostic for the tensor-core convolution kernels: compares fwd / dgrad / wgrad with
torch fp64 convolutions over a list of shapes and prints error structure.  Run on the B200 box:
    python tools/conv_diag.py [quick]
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from layout2img_b200 import ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def report(name, got, ref):
    got = got.double(); ref = ref.double()
    err = (got - ref).abs()
    scale = ref.abs().max().item() + 1e-30
    print(f"  {name:8s} max|err| {err.max().item():.3e}  rel-to-max {err.max().item()/scale:.3e}  ref max {scale:.3e}"
          f"  nan {int(torch.isnan(got).sum())}", flush=True)
    return err.max().item() / scale


def run(N, Cin, Cout, H, k, seed=0, bias=True, detail=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, Cin, H, H, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev) if bias else None
    dy = torch.randn(N, Cout, H, H, generator=g).to(dev)
    print(f"conv N={N} Cin={Cin} Cout={Cout} H={H} k={k}", flush=True)
    ref = F.conv2d(x.double(), w.double(), b.double() if bias else None, padding=k // 2)
    ref_dx = torch.nn.grad.conv2d_input(x.shape, w.double(), dy.double(), padding=k // 2)
    ref_dw = torch.nn.grad.conv2d_weight(x.double(), w.shape, dy.double(), padding=k // 2)
    wp = ops.conv_weight_prep(w)
    xp = ops.act_split(nhwc(x))
    y, _ = ops.conv2d_fwd(xp, wp.f_hi, wp.f_lo, Cout, k * k, bias=b)
    torch.cuda.synchronize()
    worst = report("fwd", y.permute(0, 3, 1, 2), ref)
    if detail and worst > 1e-3:
        e = (y.permute(0, 3, 1, 2).double() - ref).abs()
        print("   err by cout[:16]", e.amax(dim=(0, 2, 3))[:16].tolist())
        print("   err by row[:16]", e.amax(dim=(0, 1, 3))[:16].tolist())
        print("   err by col[:16]", e.amax(dim=(0, 1, 2))[:16].tolist())
        print("   err by img[:8]", e.amax(dim=(1, 2, 3))[:8].tolist())
    dyp = ops.act_split(nhwc(dy))
    dx, _ = ops.conv2d_fwd(dyp, wp.d_hi, wp.d_lo, Cin, k * k)
    torch.cuda.synchronize()
    worst = max(worst, report("dgrad", dx.permute(0, 3, 1, 2), ref_dx))
    dw = ops.conv2d_wgrad(dyp, xp, k * k)
    torch.cuda.synchronize()
    dw_t = dw.view(Cout, k, k, Cin).permute(0, 3, 1, 2)
    ww = report("wgrad", dw_t, ref_dw)
    if detail and ww > 1e-3:
        e = (dw_t.double() - ref_dw).abs()
        print("   err by tap", e.amax(dim=(0, 1)).flatten().tolist())
        print("   err by cout[:16]", e.amax(dim=(1, 2, 3))[:16].tolist())
        print("   err by cin[:16]", e.amax(dim=(0, 2, 3))[:16].tolist())
    worst = max(worst, ww)
    # timing (fwd only)
    for _ in range(2):
        ops.conv2d_fwd(xp, wp.f_hi, wp.f_lo, Cout, k * k, bias=b)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        ops.conv2d_fwd(xp, wp.f_hi, wp.f_lo, Cout, k * k, bias=b)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    fl = 2.0 * N * H * H * Cout * Cin * k * k
    t0.record()
    for _ in range(5):
        ops.conv2d_wgrad(dyp, xp, k * k)
    t1.record(); torch.cuda.synchronize()
    ms2 = t0.elapsed_time(t1) / 5
    print(f"  fwd {ms:.3f} ms = {fl/ms/1e9:.1f} TFLOP/s ; wgrad {ms2:.3f} ms = {fl/ms2/1e9:.1f} TFLOP/s", flush=True)
    return worst


if __name__ == "__main__":
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    shapes = [(2, 64, 64, 16, 3), (2, 64, 64, 16, 1), (3, 128, 128, 8, 3), (8, 128, 256, 4, 3), (2, 8, 64, 32, 3),
              (2, 64, 3, 32, 3), (2, 256, 100, 16, 3), (2, 528, 100, 16, 3), (2, 104, 184, 8, 1)]
    if not quick:
        shapes += [(64, 512, 512, 32, 3), (64, 128, 64, 128, 3), (64, 1024, 1024, 8, 3), (512, 1024, 1024, 8, 3),
                   (64, 64, 64, 128, 3), (64, 1024, 512, 16, 1)]
    bad = 0
    for s in shapes:
        try:
            w = run(*s, detail=True)
            bad += w > 1e-4
        except Exception as e:
            import traceback; traceback.print_exc()
            bad += 1
    print("FAILED shapes:", bad)
