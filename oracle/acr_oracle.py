"""CPU ORACLE for the ACR_WSSS all-pairs attention-affinity hot path.  TEST INFRASTRUCTURE ONLY.

This file restates the reference's algorithm in plain torch-CPU fp32 (no CUDA, nothing from acr_wsss_b200).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it; the
product path never does.

Parity status: PINNED BY EXECUTION.  The reference has no tests or golden files, so the
restatement is pinned against outputs of the UNMODIFIED reference Python imported from /root/reference in
the build container (tests/golden/make_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py).

Each function cites the reference lines it follows (paths relative to the ACR_WSSS tree).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# synthetic weights: deterministic per-key values, identical for the reference model, the oracle and
# the CUDA model (so no 361 MB state_dict has to be stored as a fixture)
# --------------------------------------------------------------------------------------------
def _key_seed(key):
    h = 2166136261
    for ch in key.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def synth_state_dict(shapes, qkv_gain=4.0, seed=0):
    """shapes: {key: shape}.  trunc-normal-ish(std .02) weights, small random biases, LayerNorm near identity;
    attn.qkv.weight is scaled by `qkv_gain` so the softmax is not uniform."""
    sd = {}
    for key in sorted(shapes):
        g = torch.Generator().manual_seed((_key_seed(key) + seed) & 0x7FFFFFFF)
        shape = tuple(shapes[key])
        t = torch.randn(shape, generator=g)
        if key.endswith("norm1.weight") or key.endswith("norm2.weight") or key.endswith("norm.weight"):
            t = 1.0 + 0.05 * t
        elif key.endswith(".bias"):
            t = 0.02 * t
        else:
            t = 0.02 * t.clamp(-2, 2)
        if key.endswith("attn.qkv.weight"):
            t = t * qkv_gain
        sd[key] = t.float()
    return sd


def vit_shapes(dim=768, depth=12, num_classes=20, grid=24, scratch_in=(96, 192, 384, 768), features=256):
    """state_dict key layout of the reference ACR model (checkpoint row)."""
    s = {"pretrained.model.cls_token": (1, 1, dim), "pretrained.model.bkg_token": (1, 1, dim),
         "pretrained.model.pos_embed": (1, grid * grid + 1, dim),
         "pretrained.model.patch_embed.proj.weight": (dim, 3, 16, 16), "pretrained.model.patch_embed.proj.bias": (dim,),
         "pretrained.model.norm.weight": (dim,), "pretrained.model.norm.bias": (dim,),
         "pretrained.model.head.weight": (1000, dim), "pretrained.model.head.bias": (1000,),
         "cls_head.weight": (num_classes, dim), "cls_head.bias": (num_classes,)}
    for i in range(depth):
        p = f"pretrained.model.blocks.{i}."
        s.update({p + "norm1.weight": (dim,), p + "norm1.bias": (dim,), p + "norm2.weight": (dim,), p + "norm2.bias": (dim,),
                  p + "attn.qkv.weight": (3 * dim, dim), p + "attn.qkv.bias": (3 * dim,),
                  p + "attn.proj.weight": (dim, dim), p + "attn.proj.bias": (dim,),
                  p + "mlp.fc1.weight": (4 * dim, dim), p + "mlp.fc1.bias": (4 * dim,),
                  p + "mlp.fc2.weight": (dim, 4 * dim), p + "mlp.fc2.bias": (dim,)})
    for i, c in enumerate(scratch_in):
        s[f"scratch.layer{i + 1}_rn.weight"] = (features, c, 3, 3)
    return s


# --------------------------------------------------------------------------------------------
# (a1) attention, (a3) trunk, (a4-a6) wrapper
# --------------------------------------------------------------------------------------------
def attention_core(qkv, num_heads, scale):
    """models/vision_transformer.py:198-211 between the two Linear layers.
    qkv [B,N,3E] -> (out [B,N,E], P [B,H,N,N])."""
    B, N, E3 = qkv.shape
    C = E3 // 3
    q, k, v = qkv.reshape(B, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    attn = (q @ k.transpose(-2, -1)) * scale
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return out, attn


def attention_core_backward(qkv, num_heads, scale, d_out, g_mean=None):
    """Closed form of SURVEY section 9 ("Attention backward with the affinity term").
    Returns (d_qkv, dP) with dP = dO V^T + g_mean/H (what the reference hook stores, :192-193,209)."""
    B, N, E3 = qkv.shape
    C = E3 // 3
    H, D = num_heads, C // num_heads
    q, k, v = qkv.reshape(B, N, 3, H, D).permute(2, 0, 3, 1, 4)
    P = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
    dO = d_out.reshape(B, N, H, D).permute(0, 2, 1, 3)
    dP = dO @ v.transpose(-2, -1)
    if g_mean is not None:
        dP = dP + g_mean.unsqueeze(1) / H
    dS = P * (dP - (P * dP).sum(-1, keepdim=True))
    dq = dS @ k * scale
    dk = dS.transpose(-2, -1) @ q * scale
    dv = P.transpose(-2, -1) @ dO
    d_qkv = torch.stack([dq, dk, dv], dim=0).permute(1, 3, 0, 2, 4).reshape(B, N, E3)
    return d_qkv, dP


def resize_pos_embed(posemb, gs_h, gs_w):
    """models/vision_transformer.py:490-504 (start_index = 1)."""
    tok, grid = posemb[:, :1], posemb[0, 1:]
    gs_old = int(math.sqrt(len(grid)))
    grid = grid.reshape(1, gs_old, gs_old, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(gs_h, gs_w), mode="bilinear")
    grid = grid.permute(0, 2, 3, 1).reshape(1, gs_h * gs_w, -1)
    return torch.cat([tok, grid], dim=1)


def trunk(sd, x, num_heads=12, keep_maps=True):
    """forward_flex (vision_transformer.py:449-486) + Block.forward (:230-233) + Mlp (:158-164), LN eps 1e-6.
    Returns (layer_4 = un-normalised output of the last block, [P_l] per block)."""
    pre = "pretrained.model."
    depth = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith(pre + "blocks."))
    b, c, h, w = x.shape
    pos = resize_pos_embed(sd[pre + "pos_embed"], h // 16, w // 16)
    t = F.conv2d(x, sd[pre + "patch_embed.proj.weight"], sd[pre + "patch_embed.proj.bias"], stride=16)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd[pre + "cls_token"].expand(b, -1, -1), t), dim=1) + pos
    dim = t.shape[-1]
    scale = (dim // num_heads) ** -0.5
    maps = []
    for i in range(depth):
        p = f"{pre}blocks.{i}."
        y = F.layer_norm(t, (dim,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
        qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
        o, P = attention_core(qkv, num_heads, scale)
        if keep_maps:
            if P.requires_grad:
                P.retain_grad()
            maps.append(P)
        t = t + F.linear(o, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        y = F.layer_norm(t, (dim,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
        y = F.linear(F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])),
                     sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        t = t + y
    return t, maps


def forward_cls(sd, x, num_heads=12):
    """DPT.forward_cls, DPT/ACR.py:92-116."""
    layer_4, maps = trunk(sd, x, num_heads)
    x_cls = F.linear(layer_4[:, 0, :], sd["cls_head.weight"], sd["cls_head.bias"])
    x_patch_cls = F.linear(layer_4[:, 1:, :].mean(dim=1), sd["cls_head.weight"], sd["cls_head.bias"])
    attn = torch.stack([m.mean(dim=1) for m in maps], dim=1)
    return x_cls, x_patch_cls, attn, None, maps


def forward_cam(sd, x, num_heads=12):
    """DPT.forward_cam, DPT/ACR.py:118-143."""
    layer_4, maps = trunk(sd, x, num_heads)
    x_cls = F.linear(layer_4[:, 0, :], sd["cls_head.weight"], sd["cls_head.bias"])
    x_patch_cls = F.linear(layer_4[:, 1:, :].mean(dim=1), sd["cls_head.weight"], sd["cls_head.bias"])
    x_patch_cam = F.relu(F.linear(layer_4[:, 1:, :], sd["cls_head.weight"], sd["cls_head.bias"]))
    attn = torch.stack([m.mean(dim=1) for m in maps], dim=1)
    return x_cls, x_patch_cls, attn, x_patch_cam, maps


# --------------------------------------------------------------------------------------------
# (a7) consistency loss
# --------------------------------------------------------------------------------------------
def consistency_loss_inplace(attn1, attn2, p):
    """Literal restatement of train_acr.py:143-161 (slices + 3*p in-place flips + two l1_loss), on clones
    so the caller's tensors survive.  Autograd-capable.  Returns (cls_align_loss, aff_align_loss)."""
    attn2 = attn2.clone()
    attn1_cls = attn1[:, :, 0, 1:].unsqueeze(2)
    attn2_cls = attn2[:, :, 0, 1:].unsqueeze(2)
    attn1_aff = attn1[:, :, 1:, 1:]
    attn2_aff = attn2[:, :, 1:, 1:]
    for i in range(p):
        attn2_cls[:, :, :, i * p:i * p + p] = attn2_cls[:, :, :, i * p:i * p + p].flip(3)
    for i in range(p):
        attn2_aff[:, :, i * p:i * p + p, :] = attn2_aff[:, :, i * p:i * p + p, :].flip(2)
    for i in range(p):
        attn2_aff[:, :, :, i * p:i * p + p] = attn2_aff[:, :, :, i * p:i * p + p].flip(3)
    return F.l1_loss(attn1_cls, attn2_cls, reduction="mean"), F.l1_loss(attn1_aff, attn2_aff, reduction="mean")


def flip_perm(p):
    """pi(0)=0, pi(1+r*p+c) = 1+r*p+(p-1-c)  (SURVEY section 9)."""
    idx = torch.arange(p * p).reshape(p, p).flip(1).reshape(-1) + 1
    return torch.cat([torch.zeros(1, dtype=torch.long), idx])


def consistency_loss_closed_form(attn1, attn2, p, alpha_cls=1.0, alpha_aff=1.0):
    """Permutation form + analytic gradient; fp64 accumulation for the means."""
    B, L, N, _ = attn1.shape
    pi = flip_perm(p)
    a2t = attn2[:, :, pi, :][:, :, :, pi]
    d = attn1 - a2t
    cls = d[:, :, 0, 1:].abs().double().mean().float()
    aff = d[:, :, 1:, 1:].abs().double().mean().float()
    W = torch.zeros(N, N)
    W[0, 1:] = alpha_cls / (B * L * (N - 1))
    W[1:, 1:] = alpha_aff / (B * L * (N - 1) ** 2)
    g1 = torch.sign(d) * W
    g2 = (-g1)[:, :, pi, :][:, :, :, pi]      # pi is an involution
    return cls, aff, g1, g2


def total_loss(x1, x2, label, attn1, attn2, p, alpha):
    """train_acr.py:160-168."""
    cls_align, aff_align = consistency_loss_inplace(attn1, attn2, p)
    l1 = F.multilabel_soft_margin_loss(x1, label)
    l2 = F.multilabel_soft_margin_loss(x2, label)
    return l1 + l2 + cls_align * alpha + aff_align * alpha, (l1, l2, cls_align, aff_align)


def train_step_loss(sd, img, label, alpha, num_heads=12):
    """forward_mirror (DPT/ACR.py:170-174) + loss block; sd tensors may require grad."""
    img2 = img.flip(-1)                                    # transforms.RandomHorizontalFlip(p=1), train_acr.py:135
    x1, _, attn1, _, _ = forward_cls(sd, img, num_heads)
    x2, _, attn2, _, _ = forward_cls(sd, img2, num_heads)
    p = img.shape[2] // 16
    loss, parts = total_loss(x1, x2, label, attn1, attn2, p, alpha)
    return loss, parts, (attn1, attn2, x1, x2)


# --------------------------------------------------------------------------------------------
# (a8) GETAM, (a9) CAM pipeline
# --------------------------------------------------------------------------------------------
def getam(maps, grads, batch, start_layer=0, func="grad", skip=1):
    """ACR.getam, DPT/ACR.py:177-215.  maps/grads: per-block [B,H,N,N] P and dP."""
    cam_list, attn_list = [], []
    for cam, grad in zip(maps, grads):
        attn_list.append(cam.mean(dim=1))
        cam = cam[batch].reshape(-1, cam.shape[-1], cam.shape[-1])
        grad = grad[batch].reshape(-1, grad.shape[-1], grad.shape[-1])
        if func == "cam_grad_s":
            cam = (grad * cam).clamp(min=0).mean(dim=0)
            cam = cam * grad.clamp(min=0).mean(dim=0)
        elif func == "cam_grad":
            cam = (grad * cam).clamp(min=0).mean(dim=0)
        elif func == "grad":
            cam = grad.clamp(min=0).mean(dim=0)
        elif func == "grad_s":
            cam = grad.clamp(min=0).mean(dim=0)
            cam = cam * grad.clamp(min=0).mean(dim=0)
        cam_list.append(cam.unsqueeze(0))
    cam_list = cam_list[start_layer:]
    cams = torch.stack(cam_list).sum(dim=0)
    return torch.relu(cams[:, 0, skip:]), attn_list, cam_list


def affinity_refine(attn, cam, t=1, normalize=False):
    """infer_cam.py:164-165,184 (t=1, normalize=False) and its A^t / row-normalised generalisation."""
    A = attn[:, :, 1:, 1:].sum(dim=1)
    if normalize:
        A = A / A.sum(dim=-1, keepdim=True)
    sq = cam.dim() == 2
    out = cam.unsqueeze(-1) if sq else cam
    for _ in range(t):
        out = torch.matmul(A, out)
    return out.squeeze(-1) if sq else out


def infer_cam_image(sd, img, label, out_size, scales=(1,), start_layer=9, getam_func="cam_grad_s", aff=True,
                    num_heads=12, t=1, normalize=False):
    """infer_cam.py:145-215 for one image (numpy at the end like the reference).  Device-agnostic: the GPU tests also run it
    under bf16 autocast on the GPU as the "stock PyTorch bf16" comparator."""
    C = label.shape[1]
    b, c, h, w = img.shape
    rows, cols = out_size
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    cam_list, patch_cam_list = [], []
    for scale in scales:
        for hflip in (1, 2):
            cam_matrix = torch.zeros((b, C, rows, cols), device=img.device)
            inp = F.interpolate(img, size=(int(h * scale), int(w * scale)), mode="bilinear", align_corners=False)
            if hflip % 2 == 1:
                inp = inp.flip(-1)
            ph, pw = int((h * scale) // 16), int((w * scale) // 16)
            cls_pred, _, attn, patch_cam, maps = forward_cam(sd, inp, num_heads)
            patch_cam = patch_cam.permute(0, 2, 1).reshape(1, C, ph, pw)
            patch_cam = F.interpolate(patch_cam, [rows, cols], mode="bilinear", align_corners=False)[0]
            patch_cam = patch_cam.detach().float().cpu().numpy() * label[0, :].clone().view(C, 1, 1).cpu().numpy()
            if hflip % 2 == 1:
                patch_cam = np.flip(patch_cam, axis=-1)
            patch_cam_list.append(patch_cam)
            patch_aff = torch.sum(attn[:, :, 1:, 1:], dim=1)
            if normalize:
                patch_aff = patch_aff / patch_aff.sum(-1, keepdim=True)
            output = cls_pred[0, :]
            for ci in range(C):
                if label[0, ci] > 1e-5:
                    for m in maps:
                        m.grad = None
                    one_hot = torch.zeros(C, device=img.device)
                    one_hot[ci] = 1
                    torch.sum(one_hot * output).backward(retain_graph=True)
                    cam, _, _ = getam(maps, [m.grad for m in maps], 0, start_layer, getam_func)
                    if aff:
                        cam = cam.unsqueeze(2)
                        for _ in range(t):
                            cam = torch.matmul(patch_aff, cam)
                    cam = cam.reshape(ph, pw)
                    cam = F.interpolate(cam.unsqueeze(0).unsqueeze(0), (rows, cols), mode="bilinear", align_corners=True)
                    cam_matrix[0, ci, :, :] = cam.detach().float()
            cam_up = cam_matrix[0].cpu().numpy()
            if hflip % 2 == 1:
                cam_up = np.flip(cam_up, axis=2)
            cam_list.append(cam_up)
    patch_sum = np.sum(patch_cam_list, axis=0)
    patch_norm = (patch_sum - np.min(patch_sum, (1, 2), keepdims=True)) / (
        np.max(patch_sum, (1, 2), keepdims=True) - np.min(patch_sum, (1, 2), keepdims=True) + 1e-5)
    sum_cam = np.sum(cam_list, axis=0)
    norm_cam = (sum_cam - np.min(sum_cam, (1, 2), keepdims=True)) / (
        np.max(sum_cam, (1, 2), keepdims=True) - np.min(sum_cam, (1, 2), keepdims=True) + 1e-6)
    cam_dict = {ci: norm_cam[ci] for ci in range(C) if label[0, ci] > 1e-5}
    patch_cam_dict = {ci: patch_norm[ci] for ci in range(C) if label[0, ci] > 1e-5}
    return cam_dict, patch_cam_dict, norm_cam


def pseudo_label(cam_dict, num_classes, threshold):
    """evaluation.py:30-36."""
    h, w = list(cam_dict.values())[0].shape
    tensor = np.zeros((num_classes + 1, h, w), np.float32)
    for key in cam_dict.keys():
        tensor[key + 1] = cam_dict[key]
    tensor[0, :, :] = threshold
    return np.argmax(tensor, axis=0).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# (a10) PAMR -- gather restatement of the conv-based pamr.py:10-144
# --------------------------------------------------------------------------------------------
def _gather_neighbors(x, dilations, include_center):
    """x [B,K,H,W] -> [B,K,P,H,W]; replicate padding, 3x3 taps at dilation d in row-major order
    (pamr.py:18-36,51-55), centre included only for the std kernel (pamr.py:81-98)."""
    B, K, H, W = x.shape
    outs = []
    for d in dilations:
        xp = F.pad(x, [d] * 4, mode="replicate")
        for t in range(9):
            if t == 4 and not include_center:
                continue
            dy, dx = (t // 3) * d, (t % 3) * d
            outs.append(xp[:, :, dy:dy + H, dx:dx + W])
    return torch.stack(outs, dim=2)


def pamr(x, mask, num_iter=1, dilations=(1,)):
    """PAMR.forward, pamr.py:125-144."""
    mask = F.interpolate(mask, size=x.size()[-2:], mode="bilinear", align_corners=True)
    x_std = _gather_neighbors(x, dilations, True).std(2, keepdim=True)
    xn = _gather_neighbors(x, dilations, False)
    aff = -(x.unsqueeze(2) - xn).abs() / (1e-8 + 0.1 * x_std)
    aff = aff.mean(1, keepdim=True)
    aff = F.softmax(aff, 2)
    for _ in range(num_iter):
        m = _gather_neighbors(mask, dilations, False)
        mask = (m * aff).sum(2)
    return mask


# --------------------------------------------------------------------------------------------
# DenseCRF loss on top of the bilateral filter.  NOT in the reference repository (only the call shape,
# myTool.py:825-857): parity is UNPINNED for this term; the filter itself is pinned (bilateral_oracle.py).
# --------------------------------------------------------------------------------------------
def dense_crf_loss_from_filter(seg_scaled, roi_scaled, AS, weight):
    N = seg_scaled.shape[0]
    s = seg_scaled * roi_scaled
    loss = -(s * AS).sum() / N * weight
    grad = -2.0 * AS * roi_scaled / N * weight
    return loss, grad
