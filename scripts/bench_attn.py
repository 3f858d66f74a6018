"""Per-kernel timing of the fused attention entry points at the bench shapes (CUDA events, L2 flushed between iterations)."""
import sys, os, ctypes, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acr_wsss_b200 import ops, _lib

B, N, H, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) >= 5 else (8, 785, 12, 64)))
iters = 10
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B, N, 3 * H * D, device=dev, generator=g) * 1.5).to(torch.bfloat16)
d_out = torch.randn(B, N, H * D, device=dev, generator=g).to(torch.bfloat16)
NP = (N + 3) // 4 * 4
G = (torch.randn(B, N, NP, device=dev, generator=g) * 0.01)[:, :, :N]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib()
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
out = torch.empty(B, N, H * D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device=dev)
mean = torch.empty(B, N, N, device=dev)
row0 = torch.empty(B, H, N, device=dev)
d_qkv = torch.empty_like(qkv)
wsb = L.acr_attn_bwd_bf16_workspace(B, N, H, D)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def timeit(fn):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3

def fwd(with_mean=True):
    rc = L.acr_attn_fwd_bf16(p(qkv), B, N, H, D, D ** -0.5, p(out), p(lse), p(mean) if with_mean else None, N * N, p(row0) if with_mean else None, st)
    assert rc == 0, _lib.last_error()
LD = (N + 127) // 128 * 128
codes = torch.tensor([0x00, 0x3F, 0xBF], dtype=torch.uint8, device=dev)[torch.randint(0, 3, (B, N, LD), device=dev, generator=g)]
def bwd(with_g=1):
    rc = L.acr_attn_bwd_bf16(p(qkv), p(out), p(lse), p(d_out), B, N, H, D, D ** -0.5, p(G) if with_g == 1 else None, G.stride(0), G.stride(1), p(codes) if with_g == 2 else None, codes.stride(0), codes.stride(1), 1e-7, 1e-9, None, p(d_qkv), None, p(ws), wsb, st)
    assert rc == 0, _lib.last_error()

f_core = 4.0 * B * H * N * N * D
res = {}
t = timeit(lambda: fwd(False)); res["fwd_flash_only_us"] = t; res["fwd_flash_TFLOPs"] = f_core / t / 1e6
t2 = timeit(lambda: fwd(True)); res["fwd_with_mean_us"] = t2; res["mean_kernel_us"] = t2 - t
t = timeit(lambda: bwd(0)); res["bwd_noG_us"] = t; res["bwd_noG_TFLOPs"] = 2 * f_core / t / 1e6
t = timeit(lambda: bwd(1)); res["bwd_withG_us"] = t; res["bwd_withG_TFLOPs"] = 2 * f_core / t / 1e6
t = timeit(lambda: bwd(2)); res["bwd_codes_us"] = t; res["bwd_codes_TFLOPs"] = 2 * f_core / t / 1e6
pp = int(round((N - 1) ** 0.5))
if pp * pp + 1 == N:
    a1 = torch.softmax(torch.randn(B, 12, N, N, device=dev, generator=g), -1); a2 = torch.softmax(torch.randn(B, 12, N, N, device=dev, generator=g), -1)
    t = timeit(lambda: ops.consistency_fwd_bwd(a1, a2, pp, 100.0, 100.0)); res["consistency_us"] = t; res["consistency_GBs"] = 16.0 * B * 12 * N * N / t / 1e3
    t = timeit(lambda: ops.consistency_codes(a1, a2, pp)); res["consistency_codes_us"] = t
print(json.dumps({k: round(v, 2) for k, v in res.items()}))
