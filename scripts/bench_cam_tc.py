"""Timing of the tensor-core CAM contractions (csrc/refine_tc.cu) against the CUDA-core kernels they replace (CUDA events, L2 flushed).
  python scripts/bench_cam_tc.py            all shapes, JSON line
  python scripts/bench_cam_tc.py once       one launch of each kernel at the cfg1 shape (for ncu)"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from acr_wsss_b200 import ops

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3

g = torch.Generator().manual_seed(0)
once = "once" in sys.argv
res = {}
shapes = [(2, 12, 785, 3, 1, False)] if once else [(2, 12, 785, 3, 1, False), (16, 12, 785, 3, 1, False), (2, 12, 785, 20, 2, True), (2, 12, 3137, 3, 2, True)]
for (B, L, N, C, t, norm) in shapes:
    attn = torch.softmax(torch.randn(B, L, N, N, generator=g), -1).to(dev)
    cam = torch.rand(B, N - 1, C, generator=g).to(dev)
    if once:
        ops.affinity_refine_tc(attn, cam, t, norm); torch.cuda.synchronize()
        break
    tag = f"refine_B{B}_N{N}_C{C}_t{t}"
    t_tc = timeit(lambda: ops.affinity_refine_tc(attn, cam, t, norm))
    t_cc = timeit(lambda: ops.affinity_apply(ops.affinity_sum(attn, norm), cam, t))
    alg = 4 * B * L * (N - 1) ** 2 + 2 * 4 * B * (N - 1) * C          # the maps once + cam in / out
    res[tag] = {"tc_us": round(t_tc, 1), "cuda_core_us": round(t_cc, 1), "alg_MB": round(alg / 1e6, 1), "tc_GBs": round(alg / t_tc / 1e3, 1)}
for (B, M, E, C) in [(2, 784, 768, 20)] if once else [(2, 784, 768, 20), (16, 784, 768, 20), (16, 784, 768, 80)]:
    tok = torch.randn(B, M + 1, E, generator=g).to(dev)
    W = (torch.randn(C, E, generator=g) * 0.05).to(dev); bias = torch.zeros(C, device=dev)
    if once:
        with torch.no_grad():
            ops.patch_cam(tok[:, 1:], W, bias)
        torch.cuda.synchronize()
        break
    with torch.no_grad():
        t_tc = timeit(lambda: ops.patch_cam(tok[:, 1:], W, bias))
        t_lib = timeit(lambda: F.relu(F.linear(tok[:, 1:], W, bias)))
    alg = 4 * B * M * E + 4 * C * E + 4 * B * M * C
    res[f"patch_cam_B{B}_C{C}"] = {"tc_us": round(t_tc, 1), "cublas_fp32_us": round(t_lib, 1), "alg_MB": round(alg / 1e6, 2), "tc_GBs": round(alg / t_tc / 1e3, 1)}
print(json.dumps(res))
