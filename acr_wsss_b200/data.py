"""GPU data path for the training batch.

Mirrors `get_data_from_chunk_v2` (myTool.py:1158-1199): per image RandomResizeLong -> flip -> normalise -> RandomCrop.  The
random draws stay on the host, IN THE REFERENCE'S RNG CALL ORDER (so a seeded run picks the same augmentations), the
pixel work is one gather kernel (`acr_augment_batch`, csrc/augment.cu) on the decoded uint8 images.  JPEG decoding itself
stays with the caller (cv2.imread in the reference; no GPU decoder in this image).

`Prefetcher` overlaps the host->device copy and the augmentation kernel of batch k+1 with the training step of batch k
(pinned staging, side stream): the reference loader is synchronous CPU code inside the step loop (train_acr.py:127-133).
"""
import random as _random

import numpy as np
import torch

from . import ops


def augment_params(shapes, crop_size, py_rng=None, np_rng=None):
    """The random decisions of get_data_from_chunk_v2 for images of the given (h, w) shapes, drawn in its order:
    np.random.uniform(0.7, 1.3) once (the unused `scale`, :1161), then per image np.random.uniform(0, 1) (flip_p, :1174),
    random.randint(min_long, max_long) (RandomResizeLong :996 with min_long = int(dim*0.9), max_long = int(dim/0.875), :1177)
    and two random.randrange calls in RandomCrop (:933-945, width first).  py_rng / np_rng default to the global `random`
    and `numpy.random` modules, exactly what the reference uses.  Returns an int32 array [B,12] for acr_augment_batch."""
    py_rng = py_rng or _random
    np_rng = np_rng or np.random
    np_rng.uniform(0.7, 1.3)
    dim = int(crop_size)
    min_long, max_long = int(dim * 0.9), int(dim / 0.875)
    out = np.zeros((len(shapes), 12), np.int32)
    for i, (h, w) in enumerate(shapes):
        flip_p = np_rng.uniform(0, 1)
        target_long = py_rng.randint(min_long, max_long)
        if w < h:
            tw, th = int(round(w * target_long / h)), target_long
        else:
            tw, th = target_long, int(round(h * target_long / w))
        ch, cw = min(dim, th), min(dim, tw)
        w_space, h_space = tw - dim, th - dim
        if w_space > 0:
            cont_left, img_left = 0, py_rng.randrange(w_space + 1)
        else:
            cont_left, img_left = py_rng.randrange(-w_space + 1), 0
        if h_space > 0:
            cont_top, img_top = 0, py_rng.randrange(h_space + 1)
        else:
            cont_top, img_top = py_rng.randrange(-h_space + 1), 0
        out[i] = (h, w, th, tw, int(flip_p > 0.5), img_top, img_left, cont_top, cont_left, ch, cw, 0)
    return out


class GpuAugment:
    """images (list of uint8 HWC RGB numpy arrays) -> (images [B,3,dim,dim] fp32 normalised, ori_images uint8) on `device`."""

    def __init__(self, crop_size, device, want_ori=True):
        self.dim = int(crop_size)
        self.device = torch.device(device)
        self.want_ori = want_ori
        self._stage = None
        self._copied = None         # event after the last host->device copy out of the staging buffer

    def _pinned(self, nbytes):
        if self._stage is None or self._stage.numel() < nbytes:
            self._stage = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        return self._stage

    def __call__(self, images, params=None, py_rng=None, np_rng=None):
        shapes = [im.shape[:2] for im in images]
        if params is None:
            params = augment_params(shapes, self.dim, py_rng, np_rng)
        sizes = [int(h) * int(w) * 3 for h, w in shapes]
        offs = np.zeros(len(images), np.int64)
        offs[1:] = np.cumsum(sizes)[:-1]
        total = int(sum(sizes))
        if self._copied is not None:
            self._copied.synchronize()          # the previous batch's copy must have left the staging buffer
        stage = self._pinned(total)
        view = stage.numpy()
        for im, o, n in zip(images, offs, sizes):
            assert im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3, "decoded RGB uint8 HWC images expected"
            view[o:o + n] = np.ascontiguousarray(im).reshape(-1)
        src = stage[:total].to(self.device, non_blocking=True)
        if self.device.type == "cuda":
            self._copied = torch.cuda.Event()
            self._copied.record(torch.cuda.current_stream(self.device))
        d_off = torch.from_numpy(offs).to(self.device, non_blocking=True)
        d_par = torch.from_numpy(np.ascontiguousarray(params, dtype=np.int32)).to(self.device, non_blocking=True)
        return ops.augment_batch(src, d_off, d_par, len(images), self.dim, self.want_ori)


class Prefetcher:
    """Wraps an iterator of (list of uint8 images, labels) and yields device batches one step ahead: the pinned copy and the
    augmentation kernel of the next batch run on a side stream while the caller trains on the current one."""

    def __init__(self, it, augment, label_device=None):
        self.it = iter(it)
        self.aug = augment
        self.stream = torch.cuda.Stream(device=augment.device)
        self.label_device = label_device or augment.device
        self._next = None
        self._load()

    def _load(self):
        try:
            images, labels = next(self.it)
        except StopIteration:
            self._next = None
            return
        with torch.cuda.stream(self.stream):
            img, ori = self.aug(images)
            lab = torch.as_tensor(labels).to(self.label_device, non_blocking=True)
        self._next = (img, ori, lab)

    def __iter__(self):
        return self

    def __next__(self):
        if self._next is None:
            raise StopIteration
        torch.cuda.current_stream(self.aug.device).wait_stream(self.stream)
        batch = self._next
        for t in batch:
            if t is not None:
                t.record_stream(torch.cuda.current_stream(self.aug.device))
        self._load()
        return batch
