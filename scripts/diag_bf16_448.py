"""Diagnostic: distance of the bf16 paths (fused, fused with ACR_FP32_STREAM=1, stock torch autocast) from the reference golden."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from helpers import load_golden, rel_err, t2n
from acr_wsss_b200 import ACR, synth
from oracle import acr_oracle as orc
dev = torch.device("cuda:0")
for name in ("train_vitb_64_g2.npz", "train_vitb_448_g2.npz"):
    g = load_golden(name)
    S, B, C = int(g["S"]), int(g["B"]), int(g["C"])
    sd = orc.synth_state_dict(orc.vit_shapes(768, 12, C), qkv_gain=float(g["qkv_gain"]))
    img = synth.images(B, S).to(dev)
    def errs(x1, x2, a1, a2):
        if "attn1" in g:
            return [rel_err(t2n(x1), g["x_cls_1"]), rel_err(t2n(x2), g["x_cls_2"]), rel_err(t2n(a1), g["attn1"]), rel_err(t2n(a2), g["attn2"])]
        return [rel_err(t2n(x1), g["x_cls_1"]), rel_err(t2n(x2), g["x_cls_2"]), rel_err(t2n(a1[:, ::5, ::97, ::7]), g["attn1_sub"]), rel_err(t2n(a2[:, ::5, ::97, ::7]), g["attn2_sub"])]
    for mode in ("bf16 stream", "fp32 stream"):
        m = ACR(C, "vitb", precision="bf16").to(dev); m.load_state_dict(sd); m.train()
        m.pretrained.model.residual = "fp32" if mode == "fp32 stream" else "bf16"
        cl, (a1, a2) = m.forward_mirror(img, img.flip(-1))
        print(name, mode, ["%.2e" % e for e in errs(cl[0], cl[1], a1, a2)])
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        _, _, (ta1, ta2, tx1, tx2) = orc.train_step_loss({k: v.to(dev) for k, v in sd.items()}, img, synth.labels(B, C).to(dev), 100.0)
    print(name, "stock autocast", ["%.2e" % e for e in errs(tx1, tx2, ta1, ta2)])

# CAM inference (configs[0]) against the reference golden
import numpy as np
from acr_wsss_b200 import infer_cam_image, pseudo_label
g = load_golden("infer_vitb_448_g2.npz")
C, S = int(g["C"]), int(g["S"])
present = [int(c) for c in g["present"]]
sd = orc.synth_state_dict(orc.vit_shapes(768, 12, C), qkv_gain=float(g["qkv_gain"]))
img = synth.images(1, S, seed=3).to(dev)
label = synth.labels(1, C, present=present).to(dev)
for mode in ("bf16 stream", "fp32 stream"):
    m = ACR(C, "vitb", precision="bf16").to(dev); m.load_state_dict(sd); m.eval()
    m.pretrained.model.residual = "fp32" if mode == "fp32 stream" else "bf16"
    a, pa, _ = infer_cam_image(m, img, label, (60, 80), start_layer=10, getam_func="grad")
    print("CAM", mode, "%.2e" % rel_err(np.stack([a[c] for c in present]), g["norm_cam"]), "%.2e" % rel_err(np.stack([pa[c] for c in present]), g["patch_norm_cam"]),
          [(pseudo_label(a, C, t / 100.0) == g[f"label_t{t}"]).mean() for t in (25, 40)])
