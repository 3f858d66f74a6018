import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, time
from acr_wsss_b200 import ACR, Trainer, synth
from oracle import acr_oracle as orc
dev = torch.device("cuda:0")
S, B, C = 64, 2, 20
sd = orc.synth_state_dict(orc.vit_shapes(768, 12, C), qkv_gain=2.0)
img, label = synth.images(B, S), synth.labels(B, C)
for graph in (True, False):
    m = ACR(C, "vitb", precision="bf16").to(dev); m.load_state_dict(sd)
    for n, p in m.named_parameters():
        if n.startswith(("pretrained.model.norm.", "pretrained.model.head.", "scratch.")) or n.endswith("bkg_token"):
            p.requires_grad_(False)
    tr = Trainer(m, lr=0.01, max_step=50, alpha=100.0, cuda_graph=graph)
    print(graph, [round(float(tr.step(img.pin_memory(), label.pin_memory())), 5) for _ in range(5)])
# microbench of the pieces at bench shapes
M, K, N = 12560, 768, 3072
x = torch.randn(M, K, device=dev, dtype=torch.bfloat16); w = torch.randn(N, K, device=dev, dtype=torch.bfloat16); b = torch.randn(N, device=dev, dtype=torch.bfloat16)
dy = torch.randn(M, N, device=dev, dtype=torch.bfloat16); g = torch.zeros(N, K, device=dev)
def t(fn, n=20):
    fn(); torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    a.record(); [fn() for _ in range(n)]; e.record(); torch.cuda.synchronize(); return a.elapsed_time(e) / n * 1e3
print("addmm", t(lambda: torch.addmm(b, x, w.t())), "linear", t(lambda: torch.nn.functional.linear(x, w, b)))
print("dx", t(lambda: dy @ w), "dw", t(lambda: dy.t() @ x), "sum f32", t(lambda: dy.sum(0, dtype=torch.float32)), "sum bf16", t(lambda: dy.sum(0)))
dw = dy.t() @ x
print("add_ mixed", t(lambda: g.add_(dw)), "cast+add", t(lambda: g.add_(dw.float())))
