"""TEST INFRASTRUCTURE ONLY (imported by tests/ alone).  numpy restatement of the per-image pixel work of
get_data_from_chunk_v2 (myTool.py:1171-1196): cv2.resize INTER_LINEAR coordinate rule (fx = (dx+0.5)*scale-0.5, floor, clamp
with zero weight; horizontal pass then vertical), np.fliplr, ImageNet normalisation, RandomCrop paste into a zero container
(myTool.py:923-955).  Pinned by tests/golden/augment_64.npz, which was produced by the reference's own functions + cv2."""
import numpy as np

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def _coords(n_dst, n_src):
    scale = float(n_src) / float(n_dst)
    fx = ((np.arange(n_dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(fx).astype(np.int64)
    f = (fx - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s[lo], f[lo] = 0, 0.0
    hi = s >= n_src - 1
    s[hi], f[hi] = n_src - 1, 0.0
    return s, np.minimum(s + 1, n_src - 1), f


def augment_image(img_u8, p, dim):
    """img_u8 [h,w,3] uint8; p = the 12-int record of data.augment_params.  Returns float32 [3,dim,dim]."""
    h, w, th, tw, flip, img_top, img_left, cont_top, cont_left, ch, cw, _ = (int(v) for v in p)
    y0, y1, fy = _coords(th, h)
    x0, x1, fx = _coords(tw, w)
    src = img_u8.astype(np.float64)
    hor = src[:, x0, :] * (1.0 - fx)[None, :, None] + src[:, x1, :] * fx[None, :, None]          # [h,tw,3]
    res = hor[y0] * (1.0 - fy)[:, None, None] + hor[y1] * fy[:, None, None]                      # [th,tw,3]
    if flip:
        res = res[:, ::-1]
    res = (res / 255.0 - np.array(MEAN)) / np.array(STD)
    out = np.zeros((dim, dim, 3), np.float32)
    out[cont_top:cont_top + ch, cont_left:cont_left + cw] = res[img_top:img_top + ch, img_left:img_left + cw]
    return out.transpose(2, 0, 1)
