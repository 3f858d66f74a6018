"""BASELINE.json configs[4]: ViT-L/16 at 512x512 (1025 tokens), all-pairs consistency over all 24 blocks, batch 4 per GPU."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acr_wsss_b200 import ACR, Trainer, synth

dev = torch.device("cuda", 0)
torch.manual_seed(0)
B, S, C = 4, 512, 20
model = ACR(C, "vitl", precision="bf16").to(dev)
for n, p in model.named_parameters():
    if n.startswith(("pretrained.model.norm.", "pretrained.model.head.", "scratch.")) or n.endswith("bkg_token"):
        p.requires_grad_(False)
tr = Trainer(model, lr=0.01, max_step=10 ** 6, alpha=100.0)
img, lab = synth.images(B, S, seed=0).to(dev), synth.labels(B, C, seed=0).to(dev)
losses = [float(tr.step(img, lab)) for _ in range(4)]
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    tr.step(img, lab)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(json.dumps({"workload": "ViT-L/16 512x512, 1025 tokens, 24 blocks, B=4, bf16", "ms_per_step": round(ms, 2), "img_per_s": round(B / ms * 1e3, 1),
                  "losses": [round(l, 4) for l in losses], "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1)}))
