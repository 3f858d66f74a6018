"""Where the time of infer_cam_batch goes (cfg1 workload, 8 images per pass): device part (graph replay), post-processing, host copy."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acr_wsss_b200 import ACR, synth, infer_cam_batch, cam as camm
dev = torch.device("cuda:0")
S, C, MB = 448, 20, 8
torch.manual_seed(0)
m = ACR(C, "vitb", precision="bf16").to(dev).eval()
imgb = torch.cat([synth.images(1, S, seed=200 + i) for i in range(MB)]).to(dev)
labb = synth.labels(1, C, present=(3, 7, 14)).to(dev).repeat(MB, 1)
kw = dict(start_layer=10, getam_func="grad", cuda_graph=True)
for _ in range(4):
    infer_cam_batch(m, imgb, labb, (S, S), **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    infer_cam_batch(m, imgb, labb, (S, S), **kw)
torch.cuda.synchronize()
print("total ms per batch", (time.perf_counter() - t0) / 5 * 1e3)
g = next(iter(m._cam_graphs.values()))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    g.graph.replay()
torch.cuda.synchronize()
print("graph replay ms per batch", (time.perf_counter() - t0) / 5 * 1e3)
for mode in ("fp32", "bf16"):
    m.pretrained.model.residual = mode
    m._cam_graphs.clear()
    for _ in range(4):
        infer_cam_batch(m, imgb, labb, (S, S), **kw)
    g = next(iter(m._cam_graphs.values()))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        g.graph.replay()
    torch.cuda.synchronize()
    print("residual", mode, "graph replay ms per batch", (time.perf_counter() - t0) / 5 * 1e3)
