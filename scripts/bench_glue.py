"""Timing of the trunk glue kernels at the bench shapes (CUDA events, L2 flushed): add+LayerNorm fwd/bwd, GELU fwd/bwd."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acr_wsss_b200 import ops

dev = torch.device("cuda:0")
M, E, F = 12560, 768, 3072
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, E, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
r = torch.randn(M, E, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
w = torch.ones(E, device=dev, requires_grad=True); b = torch.zeros(E, device=dev, requires_grad=True)
w.grad = torch.zeros_like(w); b.grad = torch.zeros_like(b)
bb = torch.zeros(E, device=dev, requires_grad=True); bb.grad = torch.zeros(E, device=dev)
ds = torch.randn(M, E, device=dev, generator=g).to(torch.bfloat16); dy = torch.randn(M, E, device=dev, generator=g).to(torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10

def timeit(fn):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(c)
    return tot / iters * 1e3

res = {}
res["add_ln_fwd_us"] = timeit(lambda: ops.add_layer_norm(x, r, w, b, 1e-6, True, bb))
s, y = ops.add_layer_norm(x, r, w, b, 1e-6, True, bb)
res["add_ln_bwd_us"] = timeit(lambda: torch.autograd.backward([s, y], [ds, dy], retain_graph=True))
f = torch.randn(M, F, device=dev, generator=g).to(torch.bfloat16)
yf = torch.empty_like(f); df = torch.empty_like(f)
import ctypes
from acr_wsss_b200 import _lib
L = _lib.lib(); p = lambda t: ctypes.c_void_p(t.data_ptr()); st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
col = torch.zeros(F, device=dev); wsb = L.acr_gelu_bwd_workspace(F); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
res["gelu_fwd_us"] = timeit(lambda: L.acr_gelu_fwd_bf16(p(f), p(yf), f.numel(), st))
res["gelu_bwd_us"] = timeit(lambda: L.acr_gelu_bwd_bf16(p(f), p(yf), p(df), M, F, p(col), 1, p(ws), wsb, st))
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
