import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sr3_b200
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_unet import build, TINY_UNET, rel

net = build(TINY_UNET, 32, 1)
sd = {k: v.clone() for k, v in net.state_dict().items()}
x = torch.randn(2, 6, 32, 32).cuda()
nl = torch.tensor([[0.3], [0.6]]).cuda()
e1 = net.denoise_fn(x, nl)
e1b = net.denoise_fn(x, nl)
e1c = net.denoise_fn(x, nl)
print("same engine repeat:", rel(e1b, e1), rel(e1c, e1), rel(e1c, e1b))
netB = build(TINY_UNET, 32, 1)
eB = netB.denoise_fn(x, nl)
eB2 = netB.denoise_fn(x, nl)
print("fresh engine same seed:", rel(eB, e1), rel(eB2, e1), rel(eB2, eB))
net2 = build(TINY_UNET, 32, 2)
e2 = net2.denoise_fn(x, nl)
net2.load_state_dict(sd, strict=True)
e3 = net2.denoise_fn(x, nl)
e4 = net2.denoise_fn(x, nl)
print("reload:", rel(e2, e1), rel(e3, e1), rel(e4, e1), rel(e4, e3))
for k, v in net2.state_dict().items():
    assert torch.equal(v, sd[k]), k
# compare against oracle fp32
from oracle import sr3_oracle as orc
cfg = orc.UNetConfig(6, 3, 64, 32, (1, 2), (16,), 1, 0.0, 32)
osd = {k[len("denoise_fn."):]: v.cpu() for k, v in sd.items() if k.startswith("denoise_fn.")}
with torch.no_grad():
    ref = orc.unet_forward(osd, cfg, x.cpu(), nl.cpu())
print("vs oracle: e1", rel(e1, ref), "e1b", rel(e1b, ref), "eB", rel(eB, ref), "e3", rel(e3, ref))
