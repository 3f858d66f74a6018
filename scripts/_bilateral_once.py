"""Two calls of the bilateral filter (N = 8, 224x224: K = 21 and K = 81) for an ncu launch list / capture."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acr_wsss_b200 import ops, synth
dev = torch.device("cuda:0")
for K in (21, 81):
    img = synth.smooth_rgb(8, 224, 224, seed=0).to(dev); ins = synth.probabilities(8, K, 224, 224, seed=0).to(dev)
    for _ in range(2):
        ops.bilateral_filter(img, ins, 15.0, 50.0)
torch.cuda.synchronize()
