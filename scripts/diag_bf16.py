"""Diagnostic: how far is a bf16-autocast trunk from the fp32 one, with this repo's fused attention vs torch SDPA?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from acr_wsss_b200 import ACR, synth, ops
from oracle import acr_oracle as orc

dev = torch.device("cuda:0")
S, B, C = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 2, 20
img = synth.images(B, S).to(dev)

def rel(a, b): return float((a.float() - b.float()).abs().max() / b.float().abs().max())

for gain in (4.0, 2.0, 1.0):
    sd = orc.synth_state_dict(orc.vit_shapes(768, 12, C), qkv_gain=gain)
    outs = {}
    for prec in ("fp32", "bf16"):
        m = ACR(C, "vitb", precision=prec).to(dev); m.load_state_dict(sd); m.train()
        x_cls, _, attn, _ = m.forward_cls(img)
        outs[prec] = (x_cls.detach(), attn.detach())
    # torch-only bf16: monkeypatch attention core with SDPA-equivalent math in bf16
    orig = ops.attention_core
    def torch_core(qkv, H, scale, mean_slot=None, state=None, precision="fp32"):
        Bq, N, E3 = qkv.shape; D = E3 // 3 // H
        q, k, v = qkv.reshape(Bq, N, 3, H, D).permute(2, 0, 3, 1, 4)
        P = ((q.float() @ k.float().transpose(-2, -1)) * scale).softmax(-1)
        out = (P.to(qkv.dtype) @ v).transpose(1, 2).reshape(Bq, N, H * D)
        mean = P.mean(1)
        if mean_slot is not None: mean_slot.copy_(mean.detach())
        return out, mean
    ops.attention_core = torch_core
    import acr_wsss_b200.model as mm
    m = ACR(C, "vitb", precision="bf16").to(dev); m.load_state_dict(sd); m.train()
    with torch.no_grad():
        pass
    x_cls, _, attn, _ = m.forward_cls(img)
    outs["torch_bf16"] = (x_cls.detach(), attn.detach())
    ops.attention_core = orig
    print(f"gain {gain}: ours bf16 vs fp32: logits {rel(outs['bf16'][0], outs['fp32'][0]):.3e} attn {rel(outs['bf16'][1], outs['fp32'][1]):.3e} | "
          f"torch bf16 vs fp32: logits {rel(outs['torch_bf16'][0], outs['fp32'][0]):.3e} attn {rel(outs['torch_bf16'][1], outs['fp32'][1]):.3e} | "
          f"ours vs torch bf16: logits {rel(outs['bf16'][0], outs['torch_bf16'][0]):.3e} attn {rel(outs['bf16'][1], outs['torch_bf16'][1]):.3e}")
    for l in (0, 5, 11):
        print(f"   layer {l}: ours {rel(outs['bf16'][1][:, l], outs['fp32'][1][:, l]):.3e} torch {rel(outs['torch_bf16'][1][:, l], outs['fp32'][1][:, l]):.3e}")
