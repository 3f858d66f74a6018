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
