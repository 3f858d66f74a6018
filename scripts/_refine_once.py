"""One call each of PAMR (cfg3), the bilateral filter (N=8, K=21, 224^2) and the consistency loss (cfg2) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acr_wsss_b200 import ops, synth, PAMR
dev = torch.device("cuda:0")
x = ((synth.smooth_rgb(1, 448, 448, seed=4) - 120.0) / 58.0).to(dev)
mask = synth.probabilities(1, 21, 28, 28, seed=4).to(dev)
PAMR(2, [1, 2, 4, 8, 12, 24])(x, mask)
img = synth.smooth_rgb(8, 224, 224, seed=0).to(dev); ins = synth.probabilities(8, 21, 224, 224, seed=0).to(dev)
ops.bilateral_filter(img, ins, 15.0, 50.0)
g = torch.Generator(device="cuda").manual_seed(0)
a1 = torch.softmax(torch.randn(8, 12, 785, 785, device=dev, generator=g), -1); a2 = torch.softmax(torch.randn(8, 12, 785, 785, device=dev, generator=g), -1)
ops.consistency_codes(a1, a2, 28)
ops.consistency_fwd_bwd(a1, a2, 28, 100.0, 100.0)
torch.cuda.synchronize()
# round 2: the tensor-core CAM contractions (cfg1 shapes) and the fused dense-CRF head (cfg4 shapes)
attn = torch.softmax(torch.randn(2, 12, 785, 785, device=dev, generator=g), -1)
cam = torch.rand(2, 784, 3, device=dev, generator=g)
ops.affinity_refine_tc(attn, cam, 1, False)
tok = torch.randn(2, 785, 768, device=dev, generator=g)
with torch.no_grad():
    ops.patch_cam(tok[:, 1:], torch.randn(20, 768, device=dev, generator=g) * 0.05, torch.zeros(20, device=dev))
z = (torch.randn(8, 784, 80, device=dev, generator=g) * 2.0).requires_grad_(True)
seg = ops.crf_head(z, 448)
seg.backward(torch.randn(seg.shape, device=dev, generator=g))
torch.cuda.synchronize()
