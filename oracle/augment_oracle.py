"""TEST INFRASTRUCTURE -- CPU oracle of the per-frame augmentation chain of the reference's training reader
(``CTCRAMReaderSequence2D._load_and_enqueue``, DataHandeling.py:262-395, and the static helpers it calls, :150-261):
contrast / brightness, random affine + elastic warp of image and segmentation, the segmentation relabelling
(``_fix_transformed_segmentation``), flips and rot90.  SURVEY 8f row 4.  Only tests/ and bench.py's CPU legs import it.

Parity status: **pinned**.  DataHandeling.py needs TensorFlow only for its queues; the arithmetic is numpy + OpenCV +
SciPy in static methods, which ``tests/golden/make_augment_golden.py`` imports from /root/reference (with stand-in
``tensorflow`` / ``utils`` modules) and executes as they stand on seeded inputs (tests/golden/augment.npz).  Below, the
library routines those methods call (cv2.warpAffine, scipy map_coordinates / gaussian_filter / grey_dilation) are restated
from first principles -- bit for bit, see tests/test_augment_oracle.py -- because that is what the CUDA kernels implement.
"""
import numpy as np

AB_BITS, INTER_BITS = 10, 5            # OpenCV's fixed-point affine coordinates: 1/1024 pixel, interpolation table 1/32
AB_SCALE, INTER_TAB = 1 << AB_BITS, 1 << INTER_BITS


def invert_affine(M):
    """cv2.warpAffine inverts the 2x3 matrix (no WARP_INVERSE_MAP flag) in float64 like this."""
    M = np.array(M, dtype=np.float64).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0] = A11
    M[0, 1] *= -D
    M[1, 0] *= -D
    M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2], M[1, 2] = b1, b2
    return M


def _lrint(v):
    return np.rint(v).astype(np.int64)          # cv::saturate_cast<int>(double): round half to even


def _reflect101(i, n):
    if n == 1:
        return np.zeros_like(i)
    p = 2 * (n - 1)
    i = np.mod(i, p)
    return np.where(i >= n, p - i, i)


def warp_affine_linear(img, M):
    """cv2.warpAffine(img, M, (W, H), borderMode=cv2.BORDER_REFLECT_101) for float32 images (DataHandeling.py:176)."""
    H, W = img.shape
    Mi = invert_affine(M)
    x = np.arange(W)
    adelta, bdelta = _lrint(Mi[0, 0] * x * AB_SCALE), _lrint(Mi[1, 0] * x * AB_SCALE)
    rd = AB_SCALE // INTER_TAB // 2
    out = np.zeros((H, W), np.float32)
    one = np.float32(1)
    for y in range(H):
        X0 = _lrint((Mi[0, 1] * y + Mi[0, 2]) * AB_SCALE) + rd
        Y0 = _lrint((Mi[1, 1] * y + Mi[1, 2]) * AB_SCALE) + rd
        X, Y = (X0 + adelta) >> (AB_BITS - INTER_BITS), (Y0 + bdelta) >> (AB_BITS - INTER_BITS)
        sx, sy = np.clip(X >> INTER_BITS, -32768, 32767), np.clip(Y >> INTER_BITS, -32768, 32767)
        fx = (X & (INTER_TAB - 1)).astype(np.float32) / np.float32(INTER_TAB)
        fy = (Y & (INTER_TAB - 1)).astype(np.float32) / np.float32(INTER_TAB)
        x0, x1, y0, y1 = _reflect101(sx, W), _reflect101(sx + 1, W), _reflect101(sy, H), _reflect101(sy + 1, H)
        out[y] = (img[y0, x0] * ((one - fy) * (one - fx)) + img[y0, x1] * ((one - fy) * fx)
                  + img[y1, x0] * (fy * (one - fx)) + img[y1, x1] * (fy * fx))
    return out


def warp_affine_nearest(img, M, border=-1.0):
    """cv2.warpAffine(..., borderMode=BORDER_CONSTANT, borderValue=-1, flags=INTER_NEAREST) (DataHandeling.py:172-173)."""
    H, W = img.shape
    Mi = invert_affine(M)
    x = np.arange(W)
    adelta, bdelta = _lrint(Mi[0, 0] * x * AB_SCALE), _lrint(Mi[1, 0] * x * AB_SCALE)
    rd = AB_SCALE // 2
    out = np.zeros((H, W), np.float32)
    for y in range(H):
        X = (_lrint((Mi[0, 1] * y + Mi[0, 2]) * AB_SCALE) + rd + adelta) >> AB_BITS
        Y = (_lrint((Mi[1, 1] * y + Mi[1, 2]) * AB_SCALE) + rd + bdelta) >> AB_BITS
        ok = (X >= 0) & (X < W) & (Y >= 0) & (Y < H)
        out[y] = np.where(ok, img[np.clip(Y, 0, H - 1), np.clip(X, 0, W - 1)], np.float32(border))
    return out


def _reflect_coord(c, n):
    """scipy map_coordinate(), NI_EXTEND_REFLECT (half-sample symmetric: d c b a | a b c d | d c b a)."""
    c = np.array(c, dtype=np.float64)
    s2 = 2 * n
    neg = c < 0
    v = c[neg]
    v = np.where(v < -s2, s2 * np.trunc(-v / s2) + v, v)
    c[neg] = np.where(v < -n, v + s2, -v - 1)
    pos = c > n - 1
    v = c[pos]
    v = v - s2 * np.trunc(v / s2)
    c[pos] = np.where(v >= n, s2 - v - 1, v)
    return c


def _reflect_index(i, n):
    i = np.array(i, dtype=np.int64)
    s2 = 2 * n
    neg = i < 0
    v = i[neg]
    v = np.where(v < -s2, s2 * ((-v) // s2) + v, v)
    i[neg] = np.where(v < -n, v + s2, -v - 1)
    pos = i >= n
    v = i[pos]
    v = v - s2 * (v // s2)
    i[pos] = np.where(v >= n, s2 - v - 1, v)
    return i


def map_linear_reflect(img, cy, cx):
    """scipy.ndimage.map_coordinates(img, (cy, cx), order=1, mode='reflect') -> float32 (DataHandeling.py:177)."""
    H, W = img.shape
    d = img.astype(np.float64)
    cy, cx = _reflect_coord(np.ravel(cy), H), _reflect_coord(np.ravel(cx), W)
    fy, fx = np.floor(cy), np.floor(cx)
    ty, tx = cy - fy, cx - fx
    y0, y1 = _reflect_index(fy.astype(np.int64), H), _reflect_index(fy.astype(np.int64) + 1, H)
    x0, x1 = _reflect_index(fx.astype(np.int64), W), _reflect_index(fx.astype(np.int64) + 1, W)
    t = d[y0, x0] * (1 - ty) * (1 - tx)
    t = t + d[y0, x1] * (1 - ty) * tx
    t = t + d[y1, x0] * ty * (1 - tx)
    t = t + d[y1, x1] * ty * tx
    return t.astype(np.float32).reshape(H, W)


def map_nearest_constant(img, cy, cx, cval=-1.0):
    """scipy.ndimage.map_coordinates(img, (cy, cx), order=0, mode='constant', cval=-1) (DataHandeling.py:174)."""
    H, W = img.shape
    cy, cx = np.ravel(cy).astype(np.float64), np.ravel(cx).astype(np.float64)
    ok = (cy >= 0) & (cy <= H - 1) & (cx >= 0) & (cx <= W - 1)
    iy, ix = np.floor(cy + 0.5).astype(np.int64), np.floor(cx + 0.5).astype(np.int64)
    v = np.where(ok, img[np.clip(iy, 0, H - 1), np.clip(ix, 0, W - 1)], np.float32(cval))
    return v.astype(np.float32).reshape(H, W)


def gaussian_filter_reflect(a, sigma, truncate=4.0):
    """scipy.ndimage.gaussian_filter(a, sigma) (mode='reflect'): two passes of correlate1d with the symmetric-kernel
    summation order of ni_filters.c (centre tap first, then pairs from the outermost inwards)."""
    a = np.asarray(a, dtype=np.float64)
    lw = int(truncate * float(sigma) + 0.5)
    xs = np.arange(-lw, lw + 1)
    w = np.exp(-0.5 / (float(sigma) * float(sigma)) * xs ** 2)
    w = w / w.sum()
    for axis in (0, 1):
        n = a.shape[axis]
        src = np.moveaxis(a, axis, 0)
        idx = _reflect_index(np.arange(-lw, n + lw), n)
        ext = src[idx]
        out = ext[lw:lw + n] * w[lw]
        for j in range(-lw, 0):
            out = out + (ext[lw + j:lw + j + n] + ext[lw - j:lw - j + n]) * w[j + lw]
        a = np.moveaxis(out, 0, axis)
    return a


def elastic_coords(rand2, alpha, sigma):
    """_get_indices4elastic_transform (DataHandeling.py:183-193) from the two uniform [0,1) fields the reference draws
    (first the x field, then the y field): returns (2, H, W) float64 = (y + dy, x + dx)."""
    H, W = rand2.shape[1:]
    dx = gaussian_filter_reflect(rand2[0] * 2 - 1, sigma) * alpha
    dy = gaussian_filter_reflect(rand2[1] * 2 - 1, sigma) * alpha
    x, y = np.meshgrid(np.arange(W), np.arange(H))
    return np.stack([y + dy, x + dx])


def fix_transformed_segmentation(seg):
    """_fix_transformed_segmentation (DataHandeling.py:199-211): instance labels -> {0 bg, 1 cell, 2 touching edge}."""
    r = np.round(seg)
    ri = r.astype(np.int32)
    p = np.pad(ri, 1, mode='symmetric')
    H, W = ri.shape
    dil = np.max([p[dy:dy + H, dx:dx + W] for dy in range(3) for dx in range(3)], axis=0)
    bw = np.minimum(r, 1)
    bw[(r != dil) & (dil > 0)] = 2
    return bw


def augment_frame(img, seg, contrast, brightness, affine=None, coords=None, flip=(0, 0), rot90=0, randomize=True):
    """One frame through DataHandeling.py:330-380.  img, seg float32 (H, W); contrast / brightness float32 scalars;
    affine 2x3 float64 + coords (2,H,W) float64 when elastic augmentation is on."""
    img = np.array(img, dtype=np.float32, copy=True)
    seg = np.array(seg, dtype=np.float32, copy=True)
    if randomize:
        m = img.mean()
        img = (img - m) * np.float32(contrast) + m
        img = img + np.float32(brightness)
    if affine is not None:
        img = map_linear_reflect(warp_affine_linear(img, affine), coords[0], coords[1])
        if not np.equal(seg, -1).all():
            not_valid = np.equal(seg, -1).astype(np.float32)
            seg[:, 0] = 0
            seg[:, -1] = 0
            seg[-1, :] = 0
            seg[0, :] = 0
            t_seg = map_nearest_constant(warp_affine_nearest(seg, affine), coords[0], coords[1])
            t_nv = map_nearest_constant(warp_affine_nearest(not_valid, affine), coords[0], coords[1])
            fixed = fix_transformed_segmentation(t_seg)
            fixed[(t_nv > 0.5) | (t_seg == -1)] = -1
            seg = fixed
    else:
        seg = fix_transformed_segmentation(seg)
    if flip[0]:
        img, seg = img[::-1], seg[::-1]
    if flip[1]:
        img, seg = img[:, ::-1], seg[:, ::-1]
    if rot90:
        img, seg = np.rot90(img, rot90), np.rot90(seg, rot90)
    return np.ascontiguousarray(img), np.ascontiguousarray(seg)


def synthetic_sequence(T, H, W, seed, unlabeled_every=0):
    """T frames of a cell-like image (float32, ~[0, 1000]) with an instance-labelled segmentation (float32: 0 bg,
    1.. instances, -1 = pixel without annotation; frames without any annotation are all -1)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    n = max(3, H * W // 350)
    cy, cx, r = rng.uniform(0, H, n), rng.uniform(0, W, n), rng.uniform(2, 6, n)
    imgs, segs = [], []
    for t in range(T):
        img = rng.normal(300, 30, (H, W))
        seg = np.zeros((H, W), np.float32)
        for k in range(n):
            d = np.sqrt((yy - cy[k] - 0.7 * t) ** 2 + (xx - cx[k] + 0.4 * t) ** 2)
            img += 400 * np.exp(-(d / r[k]) ** 2)
            seg[d < r[k]] = k + 1
        seg[rng.random((H, W)) < 0.01] = -1
        if unlabeled_every and t % unlabeled_every == unlabeled_every - 1:
            seg[:] = -1
        imgs.append(img.astype(np.float32))
        segs.append(seg)
    return np.stack(imgs), np.stack(segs)
