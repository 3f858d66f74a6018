"""Per-aten-op / per-kernel breakdown of one eager training step (torch.profiler, CUDA activity).  Diagnostic only:
numbers taken under a profiler are never bench values.  Usage: python scripts/prof_ops.py [--batch 8]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from torch.profiler import profile, ProfilerActivity

from acr_wsss_b200 import ACR, Trainer, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--rows", type=int, default=45)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = ACR(20, "vitb", precision="bf16").to(dev)
    for n, p in model.named_parameters():
        if n.startswith(("pretrained.model.norm.", "pretrained.model.head.", "scratch.")) or n.endswith("bkg_token"):
            p.requires_grad_(False)
    trainer = Trainer(model, lr=0.01, max_step=10 ** 6, alpha=100.0)
    img = synth.images(args.batch, 448, seed=0).to(dev)
    lab = synth.labels(args.batch, 20, seed=0).to(dev)
    for _ in range(4):
        trainer.step(img, lab)
    torch.cuda.synchronize()
    side = trainer._side
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            trainer._forward_backward(img, lab)
        torch.cuda.current_stream().wait_stream(side)
        trainer.opt.update()
        torch.cuda.synchronize()
    print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=args.rows, max_name_column_width=48,
                                                              max_shapes_column_width=70))
    print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=30, max_name_column_width=60))


if __name__ == "__main__":
    main()
