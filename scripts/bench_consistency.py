"""Timing of the consistency-loss entry (sign-code mode, the one the training step uses) at cfg2: CUDA events, L2 flushed."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acr_wsss_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
g = torch.Generator(device="cuda").manual_seed(0)
B, L, N = 8, 12, 785
a1 = torch.softmax(torch.randn(B, L, N, N, device=dev, generator=g), -1); a2 = torch.softmax(torch.randn(B, L, N, N, device=dev, generator=g), -1)
def timeit(fn, iters=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3
t_codes = timeit(lambda: ops.consistency_codes(a1, a2, 28))
t_dense = timeit(lambda: ops.consistency_fwd_bwd(a1, a2, 28, 100.0, 100.0))
print(json.dumps({"codes_us": round(t_codes, 1), "codes_GBs": round(10 * B * L * N * N / t_codes / 1e3, 1), "dense_us": round(t_dense, 1),
                  "dense_GBs": round(16 * B * L * N * N / t_dense / 1e3, 1)}))
