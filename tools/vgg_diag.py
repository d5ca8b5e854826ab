"""Diagnostic: per-tap error of vgg_loss.Vgg19 features and of the input gradient of each tap's L1 term against
torchvision VGG19 (same random weights) in fp64.  python tools/vgg_diag.py"""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from layout2img_b200.vgg_loss import VGGLoss

dev = torch.device("cuda:0")
torch.manual_seed(3)
tv = torchvision.models.vgg19(weights=None).features[:30]
with torch.no_grad():
    for m in tv:
        if isinstance(m, torch.nn.Conv2d):
            m.bias.normal_(0, 0.05)
        if isinstance(m, torch.nn.ReLU):
            m.inplace = False
ours = VGGLoss()
ours.vgg.load_state_dict({"features." + k: v for k, v in tv.state_dict().items()})
ours.to(dev)
g = torch.Generator().manual_seed(4)
x = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
y = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
net = copy.deepcopy(tv).double()

def feats(t):
    out, h = [], t
    for i, m in enumerate(net):
        h = m(h)
        if i in (1, 6, 11, 20, 29):
            out.append(h)
    return out

fy = feats(y.double())
with torch.no_grad():
    oy = ours.vgg(y.to(dev))
for tap in range(5):
    xr = x.clone().double().requires_grad_()
    fx = feats(xr)
    ((fx[tap] - fy[tap].detach()).abs().mean()).backward()
    xg = x.clone().to(dev).requires_grad_()
    ox = ours.vgg(xg)
    ((ox[tap] - oy[tap]).abs().mean()).backward()
    fe = (ox[tap].permute(0, 3, 1, 2).double().cpu() - fx[tap].detach()).abs().max().item() / fx[tap].abs().max().item()
    ge = (xg.grad.double().cpu() - xr.grad).abs().max().item() / xr.grad.abs().max().item()
    gl2 = (xg.grad.double().cpu() - xr.grad).norm().item() / xr.grad.norm().item()
    print(f"tap {tap}: feature max-rel err {fe:.3e}; input-grad max-rel err {ge:.3e}, rel-L2 {gl2:.3e}")
# L1 on a smooth surrogate (sum of squares) to separate kink effects from arithmetic
for tap in (2, 4):
    xr = x.clone().double().requires_grad_()
    (feats(xr)[tap] ** 2).mean().backward()
    xg = x.clone().to(dev).requires_grad_()
    (ours.vgg(xg)[tap] ** 2).mean().backward()
    gl2 = (xg.grad.double().cpu() - xr.grad).norm().item() / xr.grad.norm().item()
    print(f"tap {tap} (mean of squares): input-grad rel-L2 {gl2:.3e}")
