"""TEST INFRASTRUCTURE -- CPU oracle of the instance-labelling post-processing that follows the model call in the
reference's ``Inference2D.inference`` (Inference2D.py:64-123).  Only tests/, __graft_entry__.smoke()
and bench.py's CPU legs may import this module; the product path (lstm_unet_b200/postprocess.py -> lu_postprocess)
never does.

Parity status: **pinned**.  Unlike the network itself (TensorFlow, not installable here), this part of the reference
is numpy + scipy.ndimage + OpenCV, all present in this image, so the reference's own statements can be executed:
``tests/golden/make_postprocess_golden.py`` runs the lines of /root/reference/Inference2D.py as they stand (read at run
time, never copied) on seeded soft-max maps and stores inputs + outputs in ``tests/golden/postprocess.npz``;
``tests/test_postprocess_oracle.py`` checks both functions below against those vectors.

Two statements of the algorithm:

* ``postprocess_frame``   -- follows the reference step by step and calls the same third-party routines it calls
  (``scipy.ndimage.binary_fill_holes``, ``cv2.connectedComponentsWithStats``, ``scipy.ndimage.distance_transform_edt``).
* ``postprocess_frame_plain`` -- the same result from first principles (plain numpy / Python loops, no scipy / cv2): it
  spells out the library behaviours the CUDA kernels must reproduce bit for bit -- OpenCV's label numbering, SciPy's
  nearest-feature tie-breaking, the hole definition -- and is what the kernels were designed from.
"""
import numpy as np


# ---------------------------------------------------------------------------------------------------------------------
# statement 1: the reference's steps with the reference's library calls
# ---------------------------------------------------------------------------------------------------------------------
def postprocess_frame(softmax, edge_dist=2, min_cell_size=10, max_cell_size=100, FOV=0, edge_thresh=0.2,
                      return_intermediate=False):
    """softmax: (3, H, W) float32, classes (background, cell, edge) -> uint16 (H, W) instance labels.
    Defaults are CTCInferenceParams' (Params.py:164-167)."""
    import cv2
    import scipy.ndimage as ndi
    sm = np.asarray(softmax, dtype=np.float32)
    # Inference2D.py:66-69  edge = p_edge >= 0.2 ; cell = argmax == 1 and not edge
    edge = sm[2] >= np.float32(edge_thresh)
    cell = (np.argmax(sm, 0) == 1) & ~edge
    # :70-71  fill the holes of the cell mask, edge pixels that became cell stop being edge
    cell = ndi.binary_fill_holes(cell)
    edge = edge & ~cell
    # :72-75  8-connected components + areas of the cell cores
    num_cells, cc, stats, _ = cv2.connectedComponentsWithStats(cell.astype(np.uint8), 8, cv2.CV_32S)
    # :77-78  every edge pixel closer than edge_dist to a cell joins the nearest cell
    dist, ind = ndi.distance_transform_edt(~cell, return_indices=True)
    labels = cc[ind[0], ind[1]] * (edge & (dist < edge_dist)) + cc
    labels = labels.astype(np.int64)
    after_edges = labels.copy()
    # :80-91  per label: holes of the label's mask get += n (sequentially, on the running label image)
    for n in range(1, num_cells):
        bw = labels == n
        if not bw.any():
            continue
        # utils.py:51-69: the fill runs on the bounding box grown by 10 pixels (the box always keeps a ring of
        # non-label pixels or the frame border around the label, so the holes are those of the full frame)
        rows, cols = np.flatnonzero(bw.any(1)), np.flatnonzero(bw.any(0))
        r0, r1 = max(0, rows[0] - 10), min(bw.shape[0], rows[-1] + 10)
        c0, c1 = max(0, cols[0] - 10), min(bw.shape[1], cols[-1] + 10)
        box = bw[r0:r1, c0:c1]
        holes = np.zeros_like(bw)
        holes[r0:r1, c0:c1] = ndi.binary_fill_holes(box) & ~box
        labels = labels + holes * n
    # :94-104  labels without a pixel inside the field of view are dropped (the reference zeroes ONE column on the
    # left side: fov_im[:, FOV] = 0 -- kept)
    if FOV:
        fov = np.ones(labels.shape, dtype=bool)
        fov[:FOV, :] = False
        fov[-FOV:, :] = False
        fov[:, FOV] = False
        fov[:, -FOV:] = False
        present = np.unique(labels[fov])
        removed = np.setdiff1d(np.arange(num_cells), present)
    else:
        removed = np.zeros(0, dtype=np.int64)
    # :114-124  size filter on the CORE area, consecutive renumbering
    out = np.zeros(labels.shape, dtype=np.uint16)
    p = 0
    for n in range(1, num_cells):
        area = stats[n, cv2.CC_STAT_AREA]
        if min_cell_size <= area <= max_cell_size and n not in removed:
            p += 1
            out[labels == n] = p
    if return_intermediate:
        return out, {'num_cells': num_cells, 'cc': cc, 'after_edges': after_edges, 'labels': labels,
                     'area': stats[:, cv2.CC_STAT_AREA].copy(), 'kept': p}
    return out


# ---------------------------------------------------------------------------------------------------------------------
# statement 2: first principles
# ---------------------------------------------------------------------------------------------------------------------
def plain_fill_holes(mask):
    """scipy.ndimage.binary_fill_holes with its default (cross) structure: a hole is a set of False pixels that cannot
    reach the outside of the array through 4-connected False pixels."""
    mask = np.asarray(mask, dtype=bool)
    H, W = mask.shape
    reach = np.zeros((H, W), dtype=bool)
    stack = [(y, x) for y in range(H) for x in (0, W - 1) if not mask[y, x]]
    stack += [(y, x) for x in range(W) for y in (0, H - 1) if not mask[y, x]]
    for y, x in stack:
        reach[y, x] = True
    while stack:
        y, x = stack.pop()
        for yy, xx in ((y - 1, x), (y + 1, x), (y, x - 1), (y, x + 1)):
            if 0 <= yy < H and 0 <= xx < W and not mask[yy, xx] and not reach[yy, xx]:
                reach[yy, xx] = True
                stack.append((yy, xx))
    return ~reach


def plain_components8(mask):
    """cv2.connectedComponentsWithStats(mask, 8, CV_32S) -> (count incl. background, labels, areas).

    OpenCV's 8-way labelling scans 2x2 blocks in raster order and finally renumbers the surviving (smallest)
    provisional label of every component consecutively, so the components end up numbered by the raster position of
    the first 2x2 block that contains one of their pixels (key = (y // 2) * ceil(W / 2) + x // 2).  All pixels of a 2x2
    block are mutually 8-adjacent, hence keys are unique per component."""
    mask = np.asarray(mask, dtype=bool)
    H, W = mask.shape
    comp = -np.ones((H, W), dtype=np.int64)
    keys, areas = [], []
    wb = (W + 1) // 2
    for y0 in range(H):
        for x0 in range(W):
            if not mask[y0, x0] or comp[y0, x0] >= 0:
                continue
            cid = len(keys)
            comp[y0, x0] = cid
            stack = [(y0, x0)]
            key, area = 1 << 62, 0
            while stack:
                y, x = stack.pop()
                area += 1
                key = min(key, (y // 2) * wb + x // 2)
                for yy in (y - 1, y, y + 1):
                    for xx in (x - 1, x, x + 1):
                        if 0 <= yy < H and 0 <= xx < W and mask[yy, xx] and comp[yy, xx] < 0:
                            comp[yy, xx] = cid
                            stack.append((yy, xx))
            keys.append(key)
            areas.append(area)
    order = np.argsort(np.asarray(keys, dtype=np.int64), kind='stable')
    number = np.zeros(len(keys) + 1, dtype=np.int64)
    number[order + 1] = np.arange(1, len(keys) + 1)
    labels = number[comp + 1]
    area_by_label = np.zeros(len(keys) + 1, dtype=np.int64)
    for cid, a in enumerate(areas):
        area_by_label[number[cid + 1]] = a
    area_by_label[0] = int((~mask).sum())
    return len(keys) + 1, labels, area_by_label


def edge_dist_threshold(edge_dist):
    """Largest-exclusive bound on the SQUARED integer distance equivalent to ``sqrt(d2) < edge_dist`` in float64
    (the reference compares the EDT's float64 distance, Inference2D.py:78)."""
    d2 = 0
    while np.sqrt(np.float64(d2)) < np.float64(edge_dist):
        d2 += 1
    return d2            # d2' qualifies  <=>  d2' < d2


def plain_nearest_cell(cell, y, x, d2_limit):
    """Index (row, col) of the feature scipy.ndimage.distance_transform_edt(return_indices=True) reports for pixel
    (y, x), searched among cell pixels with squared distance < d2_limit; None if there is none.  SciPy's Voronoi
    feature transform resolves equidistant features towards the smallest column, then the smallest row."""
    H, W = cell.shape
    r = int(np.ceil(np.sqrt(max(d2_limit, 1))))
    best = None
    for yy in range(max(0, y - r), min(H, y + r + 1)):
        for xx in range(max(0, x - r), min(W, x + r + 1)):
            if cell[yy, xx]:
                d2 = (yy - y) ** 2 + (xx - x) ** 2
                if d2 < d2_limit:
                    cand = (d2, xx, yy)
                    if best is None or cand < best:
                        best = cand
    return None if best is None else (best[2], best[1])


def postprocess_frame_plain(softmax, edge_dist=2, min_cell_size=10, max_cell_size=100, FOV=0, edge_thresh=0.2):
    sm = np.asarray(softmax, dtype=np.float32)
    H, W = sm.shape[1:]
    edge = sm[2] >= np.float32(edge_thresh)
    # np.argmax returns the FIRST maximum
    is_cell = (sm[1] > sm[0]) & (sm[1] >= sm[2])
    nan_any = np.isnan(sm).any(0)
    if nan_any.any():                      # np.argmax treats the first NaN as the maximum
        first_nan = np.argmax(np.isnan(sm), 0)
        is_cell = np.where(nan_any, first_nan == 1, is_cell)
    cell = plain_fill_holes(is_cell & ~edge)
    edge = edge & ~cell
    num_cells, cc, area = plain_components8(cell)
    labels = cc.copy()
    lim = edge_dist_threshold(edge_dist)
    for y, x in zip(*np.nonzero(edge)):
        near = plain_nearest_cell(cell, y, x, lim)
        if near is not None:
            labels[y, x] = cc[near]
    for n in range(1, num_cells):
        bw = labels == n
        if bw.any():
            labels = labels + (plain_fill_holes(bw) & ~bw) * n
    keep = np.zeros(num_cells, dtype=bool)
    for n in range(1, num_cells):
        keep[n] = min_cell_size <= area[n] <= max_cell_size
    if FOV:
        fov = np.ones((H, W), dtype=bool)
        fov[:FOV, :] = False
        fov[-FOV:, :] = False
        fov[:, FOV] = False
        fov[:, -FOV:] = False
        inside = np.zeros(num_cells, dtype=bool)
        v = labels[fov]
        v = v[(v >= 0) & (v < num_cells)]
        inside[v] = True
        keep &= inside
    new = np.cumsum(keep) * keep
    out = np.zeros((H, W), dtype=np.uint16)
    valid = (labels >= 1) & (labels < num_cells)
    out[valid] = new[labels[valid]]
    return out


# ---------------------------------------------------------------------------------------------------------------------
# seeded synthetic soft-max maps (shared by the golden generator, the tests and bench.py)
# ---------------------------------------------------------------------------------------------------------------------
def synthetic_softmax(H, W, seed, kind='cells', n_cells=None):
    """(3, H, W) float32 soft-max map.  kind: 'cells' = elliptical cells with edge rings (some touching, some with
    holes, some nested, specks and oversize blobs); 'noise' = i.i.d. logits (adversarial: hundreds of tiny components,
    nested holes, distance ties everywhere); 'empty' = background only; 'full' = one cell covering the frame."""
    rng = np.random.default_rng(seed)
    if kind == 'empty':
        z = np.zeros((3, H, W), np.float32)
        z[0] = 4
    elif kind == 'full':
        z = np.zeros((3, H, W), np.float32)
        z[1] = 4
    elif kind == 'noise':
        z = rng.standard_normal((3, H, W)).astype(np.float32) * 1.5
        z[2] -= 1.0
    else:
        z = np.zeros((3, H, W), np.float32)
        z[0] = 2.0
        if n_cells is None:
            n_cells = max(3, H * W // 500)
        for _ in range(n_cells):
            cy, cx = rng.uniform(-2, H + 2), rng.uniform(-2, W + 2)
            a, b = rng.uniform(1.5, 6.5, 2)
            if rng.random() < 0.05:
                a, b = a * 3, b * 3                      # oversize blob
            th = rng.uniform(0, np.pi)
            ring_w = rng.uniform(0.15, 0.6)
            draws = rng.random(3)
            hole_r, hole_cls = rng.uniform(0.2, 0.5), int(rng.integers(0, 3))
            R = int(np.ceil(2.0 * max(a, b))) + 2        # everything painted below lies within r < 1.9
            y0, y1 = max(0, int(cy) - R), min(H, int(cy) + R + 1)
            x0, x1 = max(0, int(cx) - R), min(W, int(cx) + R + 1)
            if y0 >= y1 or x0 >= x1:
                continue
            yy, xx = np.mgrid[y0:y1, x0:x1].astype(np.float32)
            zz = z[:, y0:y1, x0:x1]
            u = (yy - cy) * np.cos(th) + (xx - cx) * np.sin(th)
            v = -(yy - cy) * np.sin(th) + (xx - cx) * np.cos(th)
            r = np.sqrt((u / a) ** 2 + (v / b) ** 2)
            inside = r < 1
            ring = (r >= 1) & (r < 1 + ring_w)
            zz[1][inside] = 4.0
            zz[2][inside] = np.minimum(zz[2][inside], 0)
            zz[2][ring & (zz[1] < 3)] = 3.5
            if draws[0] < 0.25:                          # a hole (background or edge class) inside the cell
                hole = r < hole_r
                zz[1][hole] = 0
                zz[hole_cls if draws[1] < 0.5 else 0][hole] = 4.5
            if draws[2] < 0.1:                           # ring of edge pixels far outside (nesting)
                far = (r >= 1.6) & (r < 1.9)
                zz[2][far] = 3.5
        z += rng.standard_normal((3, H, W)).astype(np.float32) * 0.5
    e = np.exp(z - z.max(0, keepdims=True))
    return (e / e.sum(0, keepdims=True)).astype(np.float32)
