"""TEST INFRASTRUCTURE -- CPU oracle of the per-step metrics of the reference's train / validation step
(train2D.py:97-102,111-116): the SEG measure (losses.py:29-88) and the sparse categorical accuracy.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this module.

Parity status: **pinned** for the SEG measure -- losses.py needs TensorFlow only inside ``calc_seg``; the arithmetic is
the numpy / SciPy closure ``seg_numpy``, which ``tests/golden/make_seg_golden.py`` executes as it stands (module imported
from /root/reference with a stub ``tensorflow`` entry, the closure taken from ``seg_measure(...).__closure__``) on seeded
inputs; ``tests/test_seg_oracle.py`` checks ``seg_measure`` below against those vectors (tests/golden/seg.npz).
"""
import numpy as np


def foregrounds(labels, logits, channel_axis=2, foreground_class_index=1):
    """losses.py:74-83: ground-truth foreground = (label == 1) among valid labels (> -1); predicted foreground =
    argmax over the class axis == 1.  Returns two bool arrays (B, T, H, W)."""
    gt = np.squeeze(np.asarray(labels, dtype=np.float32), channel_axis)
    gt = gt * (gt > -1).astype(np.float32)
    gt_fg = gt == foreground_class_index
    out_fg = np.argmax(np.asarray(logits), axis=channel_axis) == foreground_class_index
    return gt_fg, out_fg


def seg_measure_masks(gt_fg, out_fg):
    """losses.py:40-71 (seg_numpy): 4-connected components of both masks per frame; every ground-truth object scores
    the IoU with the predicted object covering more than half of it, else 0; mean over all objects of all frames
    (NaN if there is none).  Arithmetic in float32 like the reference (np.sum(...).astype(np.float32))."""
    import scipy.ndimage as ndi
    cross = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    scores = []
    for g_seq, s_seq in zip(gt_fg, out_fg):
        for g, s in zip(g_seq, s_seq):
            gl = ndi.label(g, structure=cross)[0]
            sl = ndi.label(s, structure=cross)[0]
            s_area = np.bincount(sl.ravel())
            for obj in range(1, gl.max() + 1):
                bw = gl == obj
                l_area = np.float32(bw.sum())
                score = 0.
                inter = np.bincount(sl[bw], minlength=1)
                for cand in np.nonzero(inter)[0]:
                    if cand == 0:
                        continue
                    i = np.float32(inter[cand])
                    if i / l_area > 0.5:
                        score = i / (l_area + np.float32(s_area[cand]) - i)
                scores.append(score)
    if not scores:
        return np.nan
    return np.mean(scores)


def seg_measure(labels, logits, channel_axis=2):
    return seg_measure_masks(*foregrounds(labels, logits, channel_axis))


def accuracy(labels, logits, channel_axis=2):
    """k.metrics.SparseCategoricalAccuracy on (label, predictions) (train2D.py:98-101): mean over ALL pixels of
    argmax(logits) == label; ignore labels (-1) never match."""
    lab = np.squeeze(np.asarray(labels, dtype=np.float32), channel_axis)
    return float(np.mean(np.argmax(np.asarray(logits), axis=channel_axis).astype(np.float32) == lab))


def synthetic_pair(B, T, H, W, seed, kind='blobs'):
    """labels (B,T,1,H,W) float32 in {-1,0,1,2} and logits (B,T,3,H,W) float32 whose foregrounds overlap partially."""
    rng = np.random.default_rng(seed)
    if kind == 'noise':
        labels = rng.integers(-1, 3, size=(B, T, 1, H, W)).astype(np.float32)
        logits = rng.standard_normal((B, T, 3, H, W)).astype(np.float32)
        return labels, logits
    labels = np.zeros((B, T, 1, H, W), np.float32)
    logits = np.zeros((B, T, 3, H, W), np.float32)
    logits[:, :, 0] = 1.0
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        for t in range(T):
            for _ in range(max(2, H * W // 400)):
                cy, cx, r = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(1.5, 6)
                d = np.sqrt((yy - cy) ** 2 + (xx - cx) ** 2)
                labels[b, t, 0][d < r] = 1
                labels[b, t, 0][(d >= r) & (d < r + 1)] = 2
                if rng.random() < 0.8:       # predicted object: shifted / shrunk copy, sometimes missing
                    dy, dx, rr = rng.uniform(-2, 2), rng.uniform(-2, 2), r * rng.uniform(0.5, 1.2)
                    d2 = np.sqrt((yy - cy - dy) ** 2 + (xx - cx - dx) ** 2)
                    logits[b, t, 1][d2 < rr] = 2.0
            labels[b, t, 0][rng.random((H, W)) < 0.02] = -1
    logits += rng.standard_normal(logits.shape).astype(np.float32) * 0.3
    return labels, logits
