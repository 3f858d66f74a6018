"""torchrun --nproc-per-node N scripts/check_sharded.py: the sharded optimiser (reduce-scatter + 1/N update + bf16 all-gather)
against the plain all-reduce + full update on the same batches: parameters after a few Trainer steps, replica identity."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from acr_wsss_b200 import ACR, Trainer, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
dev = torch.device("cuda", local)
S, B, C = 224, 2, 20
res = {}
for mode in ("0", "0b", "1"):      # the plain all-reduce twice: run-to-run noise floor (the fp32 dQ reduce-adds land in varying order)
    os.environ["ACR_SHARDED_OPT"] = mode[0]
    torch.manual_seed(0)
    m = ACR(C, "vitb", precision="bf16").to(dev)
    for n, p in m.named_parameters():
        if n.startswith(("pretrained.model.norm.", "pretrained.model.head.", "scratch.")) or n.endswith("bkg_token"):
            p.requires_grad_(False)
    tr = Trainer(m, lr=0.01, max_step=100, alpha=100.0)
    assert tr.sharded == (mode == "1")
    losses = []
    for step in range(5):
        img = synth.images(B, S, seed=10 * step + rank).to(dev)
        lab = synth.labels(B, C, seed=10 * step + rank).to(dev)
        losses.append(float(tr.step(img, lab)))
    sync = tr.check_replicas_in_sync()
    res[mode] = ({n: p.detach().float().clone() for n, p in m.named_parameters()}, losses, sync)
    del tr, m
dist_ = lambda a, b: max(float((res[a][0][n] - res[b][0][n]).abs().max() / (res[a][0][n].abs().max() + 1e-12)) for n in res[a][0])
worst, noise = dist_("0", "1"), dist_("0", "0b")
if rank == 0:
    print("losses all-reduce#2:", res["0b"][1])
    print("run-to-run noise floor (all-reduce vs all-reduce):", noise)
    print("losses all-reduce :", res["0"][1])
    print("losses sharded    :", res["1"][1])
    print("replica spread (all-reduce, sharded):", res["0"][2], res["1"][2])
    print("max relative parameter difference sharded vs all-reduce:", worst)
assert res["0"][2] == 0.0 and res["1"][2] == 0.0
assert worst <= max(3.0 * noise, 1e-4), (worst, noise)
dist.destroy_process_group()
