"""Where does the tensor-core conv's error come from?  (a) operands exactly representable in bf16
(lo halves are zero, so any error is accumulation inside the tensor core), (b) fp32 operands."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from layout2img_b200 import ops

dev = torch.device("cuda:0")
def nhwc(x): return x.permute(0, 2, 3, 1).contiguous()

def run(N, Cin, Cout, H, k, exact_bf16):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, Cin, H, H, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5
    if exact_bf16:
        x = x.bfloat16().float(); w = w.bfloat16().float()
    ref = F.conv2d(x.double().to(dev), w.double().to(dev), None, 1, k // 2)
    ref32 = F.conv2d(x.to(dev), w.to(dev), None, 1, k // 2)
    xp = ops.act_split(nhwc(x).to(dev))
    wp = ops.conv_weight_prep(w.to(dev), need_dgrad=False)
    out, _ = ops.conv2d_fwd(xp, wp.f_hi, wp.f_lo, Cout, k * k)
    out = out.permute(0, 3, 1, 2).double()
    for name, o in (("ours", out), ("torch fp32 (allow_tf32=%s)" % torch.backends.cudnn.allow_tf32, ref32.double())):
        err = o - ref
        big = ref.abs() > 0.5 * ref.abs().max()
        srel = (err * ref.sign() / ref.abs())[big].mean().item()
        print(f"  {name:32s} K={Cin*k*k:5d} exact_bf16={exact_bf16}: max|err| {err.abs().max().item():.3e} rel-to-max {err.abs().max().item()/ref.abs().max().item():.3e} "
              f"rms err/rms ref {err.pow(2).mean().sqrt().item()/ref.pow(2).mean().sqrt().item():.3e} mean signed rel err on large outputs {srel:+.3e}")

torch.backends.cudnn.allow_tf32 = False
for shape in [(8, 1024, 1024, 8, 3), (8, 512, 512, 8, 3), (8, 128, 128, 16, 3), (8, 64, 64, 16, 3), (8, 1024, 512, 8, 1)]:
    print("conv", shape)
    for ex in (True, False):
        run(*shape, ex)
