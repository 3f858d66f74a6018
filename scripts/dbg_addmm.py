import torch, time
dev = torch.device("cuda:0")
for (M, K, N) in [(768, 12560, 3072), (3072, 12560, 768), (2304, 12560, 768), (768, 12560, 768)]:
    a = torch.randn(K, M, device=dev).to(torch.bfloat16)   # dy [K rows, M]
    b = torch.randn(K, N, device=dev).to(torch.bfloat16)
    g = torch.zeros(M, N, device=dev)
    ref = (a.t().float() @ b.float())
    torch.addmm(g, a.t(), b, out_dtype=torch.float32, out=g)
    print("err", float((g - ref).abs().max() / ref.abs().max()))
    def t(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20 * 1e3
    t1 = t(lambda: torch.addmm(g, a.t(), b, out_dtype=torch.float32, out=g))
    t2 = t(lambda: g.add_(a.t() @ b))
    t3 = t(lambda: a.t() @ b)
    print(M, K, N, "addmm_f32out %.1f us   mm+add_ %.1f us   mm only %.1f us" % (t1, t2, t3))
