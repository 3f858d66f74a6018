"""Golden vectors for the GPU data path: runs the reference's OWN augmentation functions.

myTool.py cannot be imported (matplotlib / pydensecrf at import), so the unmodified source of `RandomResizeLong`, `flip`
and `RandomCrop` (myTool.py:895-899, 923-955, 995-1008) is extracted with `ast` and executed here, driven by a restatement
of the loop body of get_data_from_chunk_v2 (:1171-1196) with cv2 from this container.  Run from the repo root:
    python tests/golden/make_golden_data.py
"""
import ast
import os
import random
import sys

import cv2
import numpy as np

REF = "/root/reference/myTool.py"
src = open(REF).read()
tree = ast.parse(src)
ns = {"np": np, "cv2": cv2, "random": random}
np.bool = bool          # removed alias the reference still uses (np.bool, myTool.py:949)
np.float = float
for node in tree.body:
    if isinstance(node, ast.FunctionDef) and node.name in ("RandomResizeLong", "flip", "RandomCrop"):
        exec(compile(ast.Module([node], []), REF, "exec"), ns)

dim = 64
shapes = [(50, 70), (90, 60), (64, 64), (200, 31)]          # smaller / larger than the crop, square, very elongated
rng = np.random.RandomState(1234)
images = [rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in shapes]

random.seed(7)
np.random.seed(11)
scale = np.random.uniform(0.7, 1.3)                          # myTool.py:1161 (unused draw)
outs, oris = [], []
for im in images:
    flip_p = np.random.uniform(0, 1)
    img_temp = im.astype(np.float64)                         # cv2.cvtColor(...).astype(np.float); the input here is already RGB
    img_temp = ns["RandomResizeLong"](img_temp, int(dim * 0.9), int(dim / 0.875))
    img_temp = ns["flip"](img_temp, flip_p)
    img_temp = img_temp.copy()
    img_temp[:, :, 0] = (img_temp[:, :, 0] / 255. - 0.485) / 0.229
    img_temp[:, :, 1] = (img_temp[:, :, 1] / 255. - 0.456) / 0.224
    img_temp[:, :, 2] = (img_temp[:, :, 2] / 255. - 0.406) / 0.225
    img_temp, cropping = ns["RandomCrop"](img_temp, dim)
    ori_temp = np.zeros_like(img_temp)
    ori_temp[:, :, 0] = (img_temp[:, :, 0] * 0.229 + 0.485) * 255.
    ori_temp[:, :, 1] = (img_temp[:, :, 1] * 0.224 + 0.456) * 255.
    ori_temp[:, :, 2] = (img_temp[:, :, 2] * 0.225 + 0.406) * 255.
    outs.append(img_temp.transpose(2, 0, 1).astype(np.float32))
    oris.append(ori_temp.astype(np.uint8).transpose(2, 0, 1))

out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "augment_64.npz")
np.savez_compressed(out, dim=dim, shapes=np.array(shapes), py_seed=7, np_seed=11, img_seed=1234,
                    images=np.stack(outs), ori_images=np.stack(oris))
print("wrote", out, os.path.getsize(out), "bytes")
