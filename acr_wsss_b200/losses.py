"""Loss functions of the ACR training step.

The reference has no loss function: the block is inline at train_acr.py:140-168 (duplicated at
train_acr_coco.py:137-165).  `acr_consistency_loss` / `acr_total_loss` are the drop-ins for those lines
(SURVEY section 8b); `dense_crf_loss` is the DenseCRF regulariser whose call shape is myTool.py:825-857
(its implementation is NOT in the reference repository -- see DESIGN.md, "parity unpinned" for that term).
"""
import torch
import torch.nn.functional as F

from . import ops


def acr_consistency_loss(attn1, attn2, p, alpha=1.0):
    """Replaces train_acr.py:143-161.  attn1/attn2 [B,L,N,N] (N=p*p+1), not modified.

    Returns (cls_align_loss, aff_align_loss, weighted) where weighted = alpha*(cls+aff) carries the fused
    gradient; the two components are detached 0-dim fp32 tensors (for logging, as the reference prints them).
    """
    total, loss2 = ops.consistency_loss(attn1, attn2, p, alpha)
    return loss2[0], loss2[1], total


def acr_total_loss(x1, x2, label, attn1, attn2, p, alpha=100.0):
    """Replaces train_acr.py:143-168: BCE(x1)+BCE(x2)+alpha*(cls_align+aff_align)."""
    cls_loss_1 = F.multilabel_soft_margin_loss(x1, label)
    cls_loss_2 = F.multilabel_soft_margin_loss(x2, label)
    cls_align, aff_align, weighted = acr_consistency_loss(attn1, attn2, p, alpha)
    loss = cls_loss_1 + cls_loss_2 + weighted
    return loss, {"cls_loss_1": cls_loss_1.detach(), "cls_loss_2": cls_loss_2.detach(),
                  "cls_align_loss": cls_align, "aff_align_loss": aff_align}


class _DenseCRF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, seg, roi, sigma_rgb, sigma_xy):
        N = seg.shape[0]
        s = seg if roi is None else seg * roi               # roi None = all ones (no extra passes over the K planes)
        AS = ops.bilateral_filter(images, s, sigma_rgb, sigma_xy)
        ctx.save_for_backward(AS, roi)
        ctx.N = N
        return -torch.dot(s.reshape(-1), AS.reshape(-1)) / N

    @staticmethod
    def backward(ctx, g):
        AS, roi = ctx.saved_tensors
        gs = AS * (-2.0 * g / ctx.N)
        return None, (gs if roi is None else gs * roi), None, None, None


def dense_crf_loss(images, segmentations, rois, weight, sigma_rgb, sigma_xy, scale_factor):
    """DenseCRF regulariser on top of the bilateral filter (upstream rloss convention):
    loss = -weight/N * <S', bilateral(S')>, S' = S*ROI, all inputs down-scaled by `scale_factor`,
    sigma_xy scaled likewise; grad = -2*weight/N * bilateral(S') * ROI.
    images [N,3,H,W] in 0..255, segmentations [N,K,H,W] probabilities, rois [N,H,W]."""
    images = F.interpolate(images, scale_factor=scale_factor, recompute_scale_factor=True)
    seg = F.interpolate(segmentations, scale_factor=scale_factor, mode="bilinear", align_corners=False,
                        recompute_scale_factor=True)
    roi = F.interpolate(rois.unsqueeze(1), scale_factor=scale_factor, recompute_scale_factor=True)
    return weight * _DenseCRF.apply(images, seg, roi, sigma_rgb, sigma_xy * scale_factor)


def dense_crf_loss_from_patch_logits(images, logits, weight, sigma_rgb, sigma_xy, scale_factor=0.5):
    """dense_crf_loss(images, softmax([0, bilinear(logits -> image size)]), ones, ...) for scale_factor = 0.5 without the
    full-resolution tensors: the probabilities at half resolution come from one fused kernel (ops.crf_head) and the ROI of ones
    drops out.  images [N,3,S,S] in 0..255, logits [N,P*P,C] patch-token logits (channel-last).  Same value and gradient as the
    composition (tests/test_gpu_parity.py::test_dense_crf_from_patch_logits_matches_composition)."""
    assert scale_factor == 0.5, "the fused head implements the rloss-scale 0.5 of the COCO recipe (infer_cam.py:58-65)"
    S = images.shape[-1]
    assert images.shape[-2] == S
    images = F.interpolate(images, scale_factor=scale_factor, recompute_scale_factor=True)
    seg = ops.crf_head(logits, S)
    return weight * _DenseCRF.apply(images, seg, None, sigma_rgb, sigma_xy * scale_factor)
