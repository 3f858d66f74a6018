"""Data-path timing: GpuAugment (pinned H2D + one kernel) vs the reference's per-image CPU work restated with cv2
(myTool.py:1171-1196) on 8 decoded VOC-sized images -> 448x448 crops."""
import os, sys, time, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from acr_wsss_b200 import GpuAugment, augment_params

rng = np.random.RandomState(0)
images = [rng.randint(0, 256, size=(375, 500, 3)).astype(np.uint8) for _ in range(8)]
dim = 448
aug = GpuAugment(dim, "cuda:0")
for _ in range(3):
    aug(images)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    out, ori = aug(images)
torch.cuda.synchronize()
gpu_ms = (time.perf_counter() - t0) / 20 * 1e3
params = torch.from_numpy(augment_params([im.shape[:2] for im in images], dim)).cuda()
src = torch.from_numpy(np.concatenate([im.reshape(-1) for im in images])).cuda()
offs = torch.from_numpy(np.cumsum([0] + [im.size for im in images[:-1]]).astype(np.int64)).cuda()
from acr_wsss_b200 import ops
ops.augment_batch(src, offs, params, 8, dim)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    ops.augment_batch(src, offs, params, 8, dim)
e1.record(); torch.cuda.synchronize()
k_us = e0.elapsed_time(e1) / 50 * 1e3
cpu_ms = None
try:
    import cv2
    def cpu_one(im):
        t = random.randint(int(dim * 0.9), int(dim / 0.875))
        h, w, _ = im.shape
        shape = (int(round(w * t / h)), t) if w < h else (t, int(round(h * t / w)))
        x = cv2.resize(im.astype(np.float64), shape)
        if np.random.uniform(0, 1) > 0.5:
            x = np.fliplr(x)
        x = (x / 255.0 - np.array([0.485, 0.456, 0.406])) / np.array([0.229, 0.224, 0.225])
        c = np.zeros((dim, dim, 3), np.float32)
        hh, ww = min(dim, x.shape[0]), min(dim, x.shape[1])
        c[:hh, :ww] = x[:hh, :ww]
        return c
    t0 = time.perf_counter()
    for _ in range(5):
        [cpu_one(im) for im in images]
    cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
except ImportError:
    pass
print({"gpu_path_ms_per_batch8_incl_staging_and_h2d": round(gpu_ms, 3), "kernel_us": round(k_us, 1),
       "kernel_GBs_out": round(8 * 3 * dim * dim * 5 / k_us / 1e3, 1), "cpu_cv2_ms_per_batch8": None if cpu_ms is None else round(cpu_ms, 2)})
