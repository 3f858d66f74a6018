"""CPU tier: the N>1 host logic (gradient buckets + all-reduce, image sharding) with world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from acr_wsss_b200.parallel import GradBuckets, shard_indices
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
    unused = torch.nn.Parameter(torch.ones(3))                      # a parameter the step never touches
    params = list(net.parameters()) + [unused]
    gb = GradBuckets(params, bucket_bytes=256)                     # tiny buckets -> several all-reduces
    assert len(gb.buckets) >= 2
    x = torch.randn(5, 16, generator=torch.Generator().manual_seed(100 + rank))
    for _ in range(2):                                              # two steps: state must reset between steps
        gb.zero()
        net(x).pow(2).sum().backward()
        gb.finish()
    grads = [p.grad.clone() for p in params]
    # expected: mean over ranks of the single-process gradients
    exp = None
    for r in range(world):
        n2 = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
        n2.load_state_dict(net.state_dict())
        xr = torch.randn(5, 16, generator=torch.Generator().manual_seed(100 + r))
        n2(xr).pow(2).sum().backward()
        g = [p.grad for p in n2.parameters()]
        exp = g if exp is None else [a + b for a, b in zip(exp, g)]
    exp = [e / world for e in exp]
    ok = all(torch.allclose(a, b, atol=1e-6) for a, b in zip(grads[:-1], exp)) and float(grads[-1].abs().max()) == 0.0
    # grads are views into the flat buffer (zero-copy)
    ok = ok and all(p.grad.untyped_storage().data_ptr() == gb.flat.untyped_storage().data_ptr() for p in params)
    # sliced all-reduce + sliced optimiser update (Trainer._reduce_and_update) == whole-buffer all-reduce + one update
    from acr_wsss_b200.train import _FlatPolySGD
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8), torch.nn.Linear(8, 4))
    gb = GradBuckets(list(net.parameters()), hooks=False)          # CUDA-graph mode: no per-bucket hooks
    net(x).pow(2).sum().backward()
    local = gb.flat.clone()
    opt = _FlatPolySGD(gb.params, gb.flat, gb.offsets, lr=0.1, wt_dec=0.5, max_step=10)
    opt.set_lr_for_step()
    spans = []
    for s_, e_ in gb.reduce_chunks(3):
        spans.append((s_, e_))
        opt.update(s_, e_)
    whole = local.clone()
    dist.all_reduce(whole)
    whole /= world
    ok = ok and spans[0][0] == 0 and spans[-1][1] == gb.flat.numel() and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    ok = ok and torch.allclose(gb.flat, whole, atol=1e-6) and torch.allclose(opt.buf, whole, atol=1e-6)
    # replicas stay identical after the update
    chk = opt.flat_param.double().sum().reshape(1)
    both = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(both, chk)
    ok = ok and float((both[0] - both[1]).abs()) == 0.0
    # sharded optimiser (Trainer._reduce_and_update with self.sharded): reduce-scatter of the sharded region + all-reduce of the
    # rest, every rank updating its slice, all-gather of the bf16 copies.  The bf16 copies are complete on every rank right after
    # the step; after the sync (all-gather of the fp32 masters) the weights equal those of the plain all-reduce + full update.
    def make():
        torch.manual_seed(1)
        return torch.nn.Sequential(torch.nn.Linear(16, 40), torch.nn.ReLU(), torch.nn.Linear(40, 8), torch.nn.LayerNorm(8), torch.nn.Linear(8, 4))

    def run(sharded):
        n3 = make()
        big = [n3[0].weight, n3[0].bias, n3[2].weight] if sharded else None
        gb3 = GradBuckets(list(n3.parameters()), hooks=False, sharded=big)
        opt3 = _FlatPolySGD(gb3.params, gb3.flat, gb3.offsets, lr=0.1, wt_dec=0.5, max_step=10)
        for _ in range(3):
            gb3.zero()
            n3(x).pow(2).sum().backward()
            opt3.set_lr_for_step()
            if sharded:
                gb3.reduce_sharded()
                s_, e_ = gb3.shard_range()
                opt3.update(s_, e_, grad=gb3.grad_shard)
                opt3.update(gb3.shard_end, gb3.flat.numel())
                gb3.all_gather_shards(opt3.flat_param16)
                stale = {n: p.detach().clone() for n, p in n3.named_parameters()}
                gb3.all_gather_shards(opt3.flat_param)      # (this toy net reads the fp32 masters; the bf16 trunk reads the copies)
            else:
                gb3.reduce_all()
                opt3.update()
        p16 = {n: opt3.flat_param16[gb3.offsets[p]:gb3.offsets[p] + p.numel()].clone() for n, p in n3.named_parameters()}
        if not sharded:
            stale = None
        return {n: p.detach().clone() for n, p in n3.named_parameters()}, p16, stale, gb3

    ref_p, ref_16, _, _ = run(False)
    sh_p, sh_16, sh_stale, gbs = run(True)
    ok = ok and gbs.shard_end > 0 and gbs.shard_end % (128 * world) == 0 and gbs.shard_end < gbs.flat.numel()
    ok = ok and all(torch.allclose(ref_p[n], sh_p[n], atol=1e-6) for n in ref_p)              # fp32 masters after the sync
    ok = ok and all(torch.equal(ref_16[n], sh_16[n]) for n in ref_16)                         # bf16 copies right after the step
    ok = ok and not all(torch.allclose(ref_p[n], sh_stale[n], atol=1e-6) for n in ref_p)      # (the masters really were sharded)
    shards = shard_indices(11, rank, world)
    out[rank] = (ok, shards)
    dist.destroy_process_group()


def test_grad_buckets_allreduce_world2_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0][0] and out[1][0]
    assert sorted(out[0][1] + out[1][1]) == list(range(11)) and not set(out[0][1]) & set(out[1][1])
