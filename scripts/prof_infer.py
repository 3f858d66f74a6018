"""CPU-vs-GPU time of one CAM-inference image (torch.profiler).  Diagnostic only.
usage: prof_infer.py [bf16|fp32] [trained]   -- `trained`: run a few Trainer steps first (cached bf16 weights attached)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from torch.profiler import profile, ProfilerActivity
from acr_wsss_b200 import ACR, Trainer, synth, infer_cam_image

dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = ACR(20, "vitb", precision=sys.argv[1] if len(sys.argv) > 1 else "bf16").to(dev)
if "trained" in sys.argv:
    tr = Trainer(model, lr=0.01, max_step=1000, alpha=100.0)
    im, lb = synth.images(8, 448, seed=0).to(dev), synth.labels(8, 20, seed=0).to(dev)
    for _ in range(4):
        tr.step(im, lb)
    for p in model.parameters():
        p.grad = None
model.eval()
model.set_capture_grad(True)
img = synth.images(1, 448, seed=100).to(dev)
lab = synth.labels(1, 20, present=(3, 7, 14)).to(dev)
for _ in range(3):
    infer_cam_image(model, img, lab, (448, 448), start_layer=10, getam_func="grad")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    infer_cam_image(model, img, lab, (448, 448), start_layer=10, getam_func="grad")
torch.cuda.synchronize()
print("wall ms/image", (time.perf_counter() - t0) / 5 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    infer_cam_image(model, img, lab, (448, 448), start_layer=10, getam_func="grad")
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
