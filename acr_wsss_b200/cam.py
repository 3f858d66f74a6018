"""CAM inference with affinity refinement: the inline block infer_cam.py:145-215, lifted into functions.

`infer_cam_image` reproduces the per-image loop body (scales x flips x present classes) with the hot pieces
on this repo's kernels: head-mean stack (fused attention), GETAM from row-0 quantities, A = sum_l attn[:,l,1:,1:]
and ONE [Np x Np]x[Np x C'] contraction for all present classes instead of C' matrix-vector products.
Bilinear resizes are library calls (F.interpolate), exactly the ones the reference makes.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import ops


def affinity_refine(attn, cam, t=1, normalize=False):
    """infer_cam.py:164-165,184.  attn [B,L,N,N]; cam [B,N-1,C] or [B,N-1] -> same shape as cam.

    t=1, normalize=False is the reference (A is the plain sum over blocks, applied once);
    t>1 / normalize=True is the generalisation named by the north star (row-normalised affinity power).
    """
    squeeze = cam.dim() == 2
    c3 = cam.unsqueeze(-1) if squeeze else cam
    if c3.shape[-1] + int(bool(normalize)) <= 128:
        out = ops.affinity_refine_tc(attn, c3, t, normalize)        # tensor cores: block sum + contraction + normalisation fused
    else:                                                           # more classes than one UMMA tile: exact CUDA-core kernels
        out = ops.affinity_apply(ops.affinity_sum(attn, normalize), c3, t)
    return out.squeeze(-1) if squeeze else out


def normalize_cam(sum_cam, eps):
    """Per-class min-max normalisation, infer_cam.py:202,209 (eps 1e-5 patch CAM / 1e-6 GETAM).  [C,H,W]."""
    mn = sum_cam.amin(dim=(1, 2), keepdim=True)
    mx = sum_cam.amax(dim=(1, 2), keepdim=True)
    return (sum_cam - mn) / (mx - mn + eps)


def pseudo_label(cam_dict, num_classes, threshold):
    """evaluation.py:30-36: argmax over [threshold, cam_1..cam_C] -> uint8 label map."""
    h, w = next(iter(cam_dict.values())).shape
    tensor = np.zeros((num_classes + 1, h, w), np.float32)
    for key, v in cam_dict.items():
        tensor[key + 1] = v
    tensor[0, :, :] = threshold
    return np.argmax(tensor, axis=0).astype(np.uint8)


def save_cam_dict(path, cam_dict):
    """Pseudo-label writer, infer_cam.py:227-228: `np.save(path, {class_index: float32 [rows,cols]})` -- the pickled-dict
    .npy that evaluation.py:28-31 (and the downstream segmentation training) reads back with allow_pickle."""
    np.save(path, {int(k): np.ascontiguousarray(v, dtype=np.float32) for k, v in cam_dict.items()})


def load_cam_dict(path):
    """evaluation.py:28-29."""
    return np.load(path, allow_pickle=True).item()


def label_iou(pred_labels, gt_labels, num_cls=21):
    """evaluation.py:33-67 without its 8 worker processes: per-class IoU over a list of (prediction, ground truth) uint8
    label maps (255 = ignore) and the mean.  Returns (iou [num_cls], miou)."""
    P = np.zeros(num_cls, np.float64)
    T = np.zeros(num_cls, np.float64)
    TP = np.zeros(num_cls, np.float64)
    for predict, gt in zip(pred_labels, gt_labels):
        cal = gt < 255
        mask = (predict == gt) * cal
        for i in range(num_cls):
            P[i] += np.sum((predict == i) * cal)
            T[i] += np.sum((gt == i) * cal)
            TP[i] += np.sum((gt == i) * mask)
    iou = TP / (T + P - TP + 1e-10)
    return iou, float(np.mean(iou))


def _lowres_pass(model, both, present_idx, n_present, start_layer, getam_func, aff, t, normalize, max_replicas):
    """Device-only part of one scale of the batched path: trunk forward on both flips, blocks >= start_layer on one copy
    per present class, ONE backward, GETAM rows, affinity refinement.  Everything that depends on WHICH classes are present
    goes through the device tensor `present_idx`, so the pass is CUDA-graph capturable per (shape, n_present).
    Returns (patch_cam [2,Np,C], cams [2,Np,n_present])."""
    rows_v = [[], []]
    attn = patch_cam = None
    for c0 in range(0, n_present, max_replicas):
        n = min(max_replicas, n_present - c0)
        cls_rep, _, attn, patch_cam = model.forward_cam_batched(both, n, start_layer)
        model.backward_for_getam_batched(cls_rep, present_idx[c0:c0 + n].repeat(2))
        for v in range(2):
            for k in range(n):
                cam, _, _ = model.getam(v * n + k, start_layer=start_layer, func=getam_func)
                rows_v[v].append(cam[0])
    cams = torch.stack([torch.stack(rows_v[v], dim=1) for v in range(2)])            # [2,Np,C']
    if aff:
        cams = torch.cat([affinity_refine(attn[v:v + 1].detach(), cams[v:v + 1], t=t, normalize=normalize) for v in range(2)])
    return patch_cam.detach(), cams


def _lowres_pass_batch(model, both, idx, n, start_layer, getam_func, aff, t, normalize):
    """_lowres_pass for M images at once.  both [2M,3,h,w] (flipped views first), idx [2M*n] class of every (view, image, copy);
    n copies per image (images with fewer present classes carry dummy copies).  Returns (patch_cam [2M,Np,C], cams [2M,Np,n])."""
    cls_rep, _, attn, patch_cam = model.forward_cam_batched(both, n, start_layer)
    model.backward_for_getam_batched(cls_rep, idx)
    cams = model.getam_batch(start_layer=start_layer, func=getam_func)           # [2M*n, Np]
    cams = cams.view(both.shape[0], n, -1).transpose(1, 2).contiguous()           # [2M, Np, n]
    if aff:
        cams = affinity_refine(attn.detach(), cams, t=t, normalize=normalize)
    return patch_cam.detach(), cams


class _LowresGraph:
    """CUDA graph of _lowres_pass for one (model, input shape, number of present classes, options) key: the pass is ~270
    launches of small kernels at batch 2 and is CPU-launch bound when run eagerly (8.4 ms per image, 2.8 ms of GPU time)."""

    def __init__(self, model, shape, n_present, args, batch=False):
        dev = next(model.parameters()).device
        self.both = torch.empty(shape, device=dev)
        self.idx = torch.zeros(shape[0] * n_present if batch else n_present, device=dev, dtype=torch.long)
        self.model, self.n, self.args = model, n_present, args
        self.fn = _lowres_pass_batch if batch else _lowres_pass
        self.graph = None
        self.calls = 0
        self.stream = torch.cuda.Stream(device=dev)

    def run(self, both, present_idx):
        self.both.copy_(both)
        self.idx.copy_(present_idx)
        self.calls += 1
        cur = torch.cuda.current_stream()
        if self.graph is None and self.calls <= 2:          # warm up eagerly on the capture stream (lazy inits, cuBLAS workspaces)
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                out = self.fn(self.model, self.both, self.idx, self.n, *self.args)
            cur.wait_stream(self.stream)
            return out
        if self.graph is None:
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.out = self.fn(self.model, self.both, self.idx, self.n, *self.args)
        self.graph.replay()
        return self.out


def infer_cam_image(model, img, label, out_size, scales=(1,), start_layer=9, getam_func="cam_grad_s",
                    aff=True, t=1, normalize=False, truncate_backward=True, batch_classes=True, max_replicas=8,
                    cuda_graph=False):
    """One image of the infer_cam.py loop body (:145-215).

    img [1,3,h,w] (normalised), label [1,C] multi-hot, out_size = (rows, cols) of the original image (the
    reference calls these W,H at infer_cam.py:136).  Returns (cam_dict, patch_cam_dict, norm_cam [C,rows,cols])
    with {class_index: float32 [rows,cols]} dicts as saved by np.save at :227-228.
    truncate_backward: stop each per-class backward at block `start_layer` (identical GETAM).
    batch_classes (needs truncate_backward): both flips go through the trunk as one batch of 2, the blocks >= start_layer
    run on one copy of the token stream per present class (at most max_replicas per pass) and ONE backward delivers every
    class's attention gradients for both flips.
    cuda_graph (batched path on a GPU): replay the device-only part as a CUDA graph, cached on the model per (input shape,
    number of present classes, options); the first two calls of a key run eagerly.  In-place weight updates are picked
    up by the replays (the graph reads the parameter storage); re-allocating parameters (e.g. model.to(...)) needs
    `model._cam_graphs.clear()`.
    """
    assert img.shape[0] == 1, "the reference infers one image at a time (infer_cam.py:122)"
    C = label.shape[1]
    b, c, h, w = img.shape
    rows, cols = out_size
    lab_host = label[0].tolist()                       # one device->host read (the reference tests the labels one by one)
    present = [ci for ci in range(C) if lab_host[ci] > 1e-5]
    cam_list, patch_cam_list = [], []
    nblocks = len(model.pretrained.model.blocks)
    batched = truncate_backward and batch_classes and len(present) > 0 and 0 < start_layer < nblocks
    present_idx = torch.tensor(present, device=img.device, dtype=torch.long) if present else None

    def finish_view(attn, patch_cam, cams, flipped, ph, pw):
        """infer_cam.py:153-199 for one flip: patch CAM and (affinity-refined) GETAM maps at the original image size.
        cams [1,Np,C'] (already refined when attn is None)."""
        patch_cam = patch_cam.permute(0, 2, 1).reshape(1, C, ph, pw)
        patch_cam = F.interpolate(patch_cam, [rows, cols], mode="bilinear", align_corners=False)[0]
        patch_cam = patch_cam.detach() * label[0, :].view(C, 1, 1)
        if flipped:
            patch_cam = patch_cam.flip(-1)
        patch_cam_list.append(patch_cam)
        cam_matrix = torch.zeros(C, rows, cols, device=img.device)
        if present:
            if aff and attn is not None:
                cams = affinity_refine(attn.detach(), cams, t=t, normalize=normalize)
            cams = cams[0].t().reshape(len(present), 1, ph, pw)
            cams = F.interpolate(cams, (rows, cols), mode="bilinear", align_corners=True)[:, 0]
            cam_matrix.index_copy_(0, present_idx, cams)
        if flipped:
            cam_matrix = cam_matrix.flip(-1)
        cam_list.append(cam_matrix)

    for scale in scales:
        inp = F.interpolate(img, size=(int(h * scale), int(w * scale)), mode="bilinear", align_corners=False)
        ph, pw = int((h * scale) // 16), int((w * scale) // 16)
        if batched:
            # both flips as one batch of 2, every present class as a copy of the token stream from block start_layer on:
            # one forward and ONE backward per scale (the reference: 2 forwards and 2*C' full backwards)
            both = torch.cat([inp.flip(-1), inp], dim=0)          # hflip = 1 (flipped) first, then 2, as infer_cam.py:148-151
            args = (start_layer, getam_func, aff, t, normalize, max_replicas)
            if cuda_graph and img.is_cuda:
                cache = model.__dict__.setdefault("_cam_graphs", {})
                key = (tuple(both.shape), len(present)) + args
                if key not in cache:
                    if len(cache) >= 8:
                        cache.clear()
                    cache[key] = _LowresGraph(model, tuple(both.shape), len(present), args)
                patch_cam, cams = cache[key].run(both, present_idx)
            else:
                patch_cam, cams = _lowres_pass(model, both, present_idx, len(present), *args)
            for v in range(2):
                finish_view(None, patch_cam[v:v + 1], cams[v:v + 1], v == 0, ph, pw)
            continue
        for hflip in (1, 2):
            view = inp.flip(-1) if hflip % 2 == 1 else inp
            cls_pred, _, attn, patch_cam = model.forward_cam(view)
            output = cls_pred[0, :]
            rows0 = []
            for ci in present:
                if truncate_backward:
                    model.backward_for_getam(output[ci], start_layer)
                else:
                    # (set_to_none=False: a Trainer-attached model keeps its .grad views of the flat gradient buffer)
                    model.zero_grad(set_to_none=False)
                    output[ci].backward(retain_graph=True)      # one_hot * output, infer_cam.py:173-179
                cam, _, _ = model.getam(0, start_layer=start_layer, func=getam_func)
                rows0.append(cam[0])
            cams = torch.stack(rows0, dim=1).unsqueeze(0) if present else None        # [1,Np,C']
            finish_view(attn, patch_cam, cams, hflip % 2 == 1, ph, pw)
    patch_sum = torch.stack(patch_cam_list).sum(0)
    patch_norm = normalize_cam(patch_sum, 1e-5)
    sum_cam = torch.stack(cam_list).sum(0)
    norm_cam = normalize_cam(sum_cam, 1e-6)
    # only the present classes go to the host (what the reference keeps, infer_cam.py:217-228), through one pinned buffer
    if present:
        both = torch.stack([norm_cam.index_select(0, present_idx), patch_norm.index_select(0, present_idx)])     # [2,C',rows,cols]
        host_np = _to_host(both)
    cam_dict = {ci: host_np[0, k] for k, ci in enumerate(present)}
    patch_cam_dict = {ci: host_np[1, k] for k, ci in enumerate(present)}
    return cam_dict, patch_cam_dict, norm_cam


def _to_host(t):
    """Device tensor -> numpy array through a pinned buffer of torch's caching host allocator.  The array is a VIEW of that
    buffer and keeps it alive (no second host-side copy: np.copy of the 4.8 MB per image cost 1 ms, a third of the whole call);
    the allocator recycles the block once the caller drops the arrays."""
    if not t.is_cuda:
        return t.detach().cpu().numpy()
    host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    host.copy_(t, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def infer_cam_batch(model, imgs, labels, out_size, scales=(1,), start_layer=9, getam_func="cam_grad_s", aff=True, t=1,
                    normalize=False, cuda_graph=False):
    """infer_cam_image for M images of one shape in ONE pass per scale (the reference loops over images, infer_cam.py:122-215):
    both flips of all images go through the trunk as a batch of 2M, the blocks >= start_layer run on n copies per image
    (n = the largest number of present classes in the batch; images with fewer classes carry dummy copies that are dropped),
    one backward, one batched GETAM, one batched affinity contraction.  imgs [M,3,h,w], labels [M,C], out_size (rows, cols)
    common to the batch.  Returns a list of (cam_dict, patch_cam_dict) per image, same contents as infer_cam_image."""
    M, C = labels.shape
    assert imgs.shape[0] == M
    rows, cols = out_size
    nblocks = len(model.pretrained.model.blocks)
    assert 0 < start_layer < nblocks, "the batched path stops the backward at block start_layer"
    lab_host = labels.tolist()                                   # one device->host read
    present = [[ci for ci in range(C) if lab_host[m][ci] > 1e-5] for m in range(M)]
    n = max(1, max(len(p) for p in present))
    padded = [p + [p[0] if p else 0] * (n - len(p)) for p in present]            # dummy copies repeat a class; dropped below
    idx_m = torch.tensor(padded, device=imgs.device, dtype=torch.long)           # [M,n]
    idx = idx_m.repeat(2, 1).reshape(-1)                                          # sample (v, m), copy k -> class padded[m][k]
    b, c, h, w = imgs.shape
    sum_cam = torch.zeros(M, n, rows, cols, device=imgs.device)
    sum_patch = torch.zeros(M, n, rows, cols, device=imgs.device)
    lab_sel = labels.gather(1, idx_m).view(M, n, 1, 1)
    for scale in scales:
        inp = F.interpolate(imgs, size=(int(h * scale), int(w * scale)), mode="bilinear", align_corners=False)
        ph, pw = int((h * scale) // 16), int((w * scale) // 16)
        both = torch.cat([inp.flip(-1), inp], dim=0)             # hflip = 1 (flipped) first, then 2, as infer_cam.py:148-151
        args = (start_layer, getam_func, aff, t, normalize)
        if cuda_graph and imgs.is_cuda:
            cache = model.__dict__.setdefault("_cam_graphs", {})
            key = ("batch", tuple(both.shape), n) + args
            if key not in cache:
                if len(cache) >= 8:
                    cache.clear()
                cache[key] = _LowresGraph(model, tuple(both.shape), n, args, batch=True)
            patch_cam, cams = cache[key].run(both, idx)
        else:
            patch_cam, cams = _lowres_pass_batch(model, both, idx, n, *args)
        # patch CAM of the (padded) present classes, infer_cam.py:153-158: [2M,Np,C] -> [2,M,n,rows,cols]
        pc = patch_cam.permute(0, 2, 1).reshape(2, M, C, ph, pw).gather(2, idx_m.view(1, M, n, 1, 1).expand(2, M, n, ph, pw))
        pc = F.interpolate(pc.reshape(2 * M, n, ph, pw), [rows, cols], mode="bilinear", align_corners=False).view(2, M, n, rows, cols)
        pc = pc * lab_sel
        sum_patch += pc[0].flip(-1) + pc[1]
        # GETAM maps, infer_cam.py:185-187
        cm = cams.transpose(1, 2).reshape(2 * M * n, 1, ph, pw)
        cm = F.interpolate(cm, (rows, cols), mode="bilinear", align_corners=True).view(2, M, n, rows, cols)
        sum_cam += cm[0].flip(-1) + cm[1]
    norm = torch.stack([normalize_cam(sum_cam[m], 1e-6) for m in range(M)])      # per-class min-max over the image, :202-209
    pnorm = torch.stack([normalize_cam(sum_patch[m], 1e-5) for m in range(M)])
    both_out = torch.stack([norm, pnorm])                                         # [2,M,n,rows,cols]
    host_np = _to_host(both_out)
    return [({ci: host_np[0, m, k] for k, ci in enumerate(present[m])}, {ci: host_np[1, m, k] for k, ci in enumerate(present[m])})
            for m in range(M)]
