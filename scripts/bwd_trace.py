"""clock64 timeline of one CTA of attn_bwd_kernel (development aid).  Build the instrumented library first:
`make -C acr_wsss_b200/csrc trace`; then `python scripts/bwd_trace.py [B N H]`.  Prints, per query tile, the cycle (relative to
the CTA's start) at which each role reached its trace points."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acr_wsss_b200 import _lib

_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "micro", "libacr_b200_trace.so")
L = _lib.lib()
L.acr_bwd_trace_read.argtypes = [ctypes.POINTER(ctypes.c_ulonglong)]
L.acr_bwd_trace_read.restype = None

B, N, H = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (16, 785, 12)))
D = 64
mode = int(os.environ.get("G", "2"))      # 0 no G, 1 fp32 G, 2 codes
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B, N, 3 * H * D, device=dev, generator=g) * 1.5).to(torch.bfloat16)
d_out = torch.randn(B, N, H * D, device=dev, generator=g).to(torch.bfloat16)
out = torch.empty(B, N, H * D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, N, device=dev)
d_qkv = torch.empty_like(qkv)
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
assert L.acr_attn_fwd_bf16(p(qkv), B, N, H, D, D ** -0.5, p(out), p(lse), None, 0, None, st) == 0, _lib.last_error()
LD = (N + 127) // 128 * 128
codes = torch.tensor([0x00, 0x3F, 0xBF], dtype=torch.uint8, device=dev)[torch.randint(0, 3, (B, N, LD), device=dev, generator=g)]
NP = (N + 3) // 4 * 4
G = (torch.randn(B, N, NP, device=dev, generator=g) * 0.01)[:, :, :N]
wsb = L.acr_attn_bwd_bf16_workspace(B, N, H, D)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
for _ in range(3):
    rc = L.acr_attn_bwd_bf16(p(qkv), p(out), p(lse), p(d_out), B, N, H, D, D ** -0.5, p(G) if mode == 1 else None, G.stride(0), G.stride(1),
                             p(codes) if mode == 2 else None, codes.stride(0), codes.stride(1), 1e-7, 1e-9, None, p(d_qkv), None, p(ws), wsb, st)
    assert rc == 0, _lib.last_error()
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (4 * 16 * 8))()
L.acr_bwd_trace_read(buf)
t = [[[buf[(r * 16 + i) * 8 + e] for e in range(8)] for i in range(16)] for r in range(4)]
t0 = t[3][0][0]
rel = lambda v: (v - t0) if v else None
nt = min(16, 2 * ((N + 127) // 128))
print(f"CTA 5 (its first two work items), B={B} N={N} H={H}, G mode {mode}; cycles since CTA start; kernel end {rel(t[3][0][1])}, ")
print("tile | MMA: scores issued, wait pds_full, pds_full seen, grads issued | softmax: wait sdp, sdp seen, pds_full arrive | drain: wait dq, dq seen, staged from, reduce issued")
for i in range(nt):
    m, s, d = t[0][i], t[1][i], t[2][i]
    print(f"{i:4d} | {rel(m[0])} {rel(m[1])} {rel(m[2])} {rel(m[3])} | {rel(s[0])} {rel(s[1])} {rel(s[4])} | {rel(d[0])} {rel(d[1])} {rel(d[2])} {rel(d[3])}")
s0, s15 = t[1][0], t[1][15]
print(f"softmax tile 0: sdp_free arrive {rel(s0[2])}; last tile: sdp_free arrive {rel(s15[2])}, pds_free wait from {rel(s15[3])} to {rel(s15[5])}")
