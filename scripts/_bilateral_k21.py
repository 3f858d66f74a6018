import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acr_wsss_b200 import ops, synth
dev = torch.device("cuda:0")
img = synth.smooth_rgb(8, 224, 224, seed=0).to(dev); ins = synth.probabilities(8, 21, 224, 224, seed=0).to(dev)
for _ in range(2):
    ops.bilateral_filter(img, ins, 15.0, 50.0)
torch.cuda.synchronize()
