"""The dominant convolution launches in isolation, for `ncu --set full` (3 launches per shape after warm-up):
   ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/conv_r02 python tools/profile_conv.py
Shapes: D.block_obj5.conv2 forward (the roofline launch of bench.py), its weight gradient, and 64->64 @128^2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from layout2img_b200 import ops

dev = "cuda"
g = torch.Generator().manual_seed(0)
rt = torch.cuda.cudart()
for (N, Cin, Cout, H) in ((512, 1024, 1024, 8), (64, 64, 64, 128)):
    x = torch.randn(N, H, H, Cin, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev)
    dy = torch.randn(N, H, H, Cout, generator=g).to(dev)
    bias = torch.randn(Cout, generator=g).to(dev)
    sc = torch.randn(N, H // 2, H // 2, Cout, generator=g).to(dev)       # the pooled 1x1 shortcut, added in the epilogue
    wp = ops.conv_weight_prep(w)
    xp, dyp = ops.act_split(x), ops.act_split(dy)
    # the in-step form of a down-sampling D block's conv2: bias, 2x2 average pooling and shortcut add in the epilogue
    fwd = lambda: ops.conv2d_fwd(xp, wp.f_hi, wp.f_lo, Cout, 9, bias=bias, residual=sc, pool=1)
    for _ in range(3):
        fwd()
        ops.conv2d_wgrad(dyp, xp, 9)
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    for _ in range(2):
        fwd()
    ops.conv2d_wgrad(dyp, xp, 9)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
    del x, w, dy, wp, xp, dyp, sc, bias
