"""CPU tests of the post-processing oracle (oracle/postprocess_oracle.py): both statements against the golden vectors
produced by executing the reference's own statements (tests/golden/make_postprocess_golden.py), and the first-principles
building blocks against the libraries the reference calls (OpenCV label numbering, SciPy hole filling and
nearest-feature tie-breaking)."""
import os

import numpy as np
import pytest

from oracle import postprocess_oracle as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'postprocess.npz')


def golden_cases():
    g = np.load(GOLD)
    for name in g['names']:
        name = str(name)
        e, mn, mx, fov = g[name + '/params']
        kw = dict(edge_dist=float(e) if e != int(e) else int(e), min_cell_size=int(mn), max_cell_size=int(mx), FOV=int(fov))
        yield name, g[name + '/softmax'].astype(np.float32), g[name + '/labels'], int(g[name + '/num_cells']), kw


@pytest.mark.parametrize('case', list(golden_cases()), ids=lambda c: c[0])
def test_oracle_matches_reference_vectors(case):
    name, sm, want, num, kw = case
    got, info = P.postprocess_frame(sm, return_intermediate=True, **kw)
    assert info['num_cells'] == num
    assert got.dtype == np.uint16 and np.array_equal(got, want)


@pytest.mark.parametrize('case', [c for c in golden_cases() if c[1].shape[1] * c[1].shape[2] <= 100 * 130],
                         ids=lambda c: c[0])
def test_plain_statement_matches_reference_vectors(case):
    name, sm, want, num, kw = case
    assert np.array_equal(P.postprocess_frame_plain(sm, **kw), want)


def test_components_numbering_is_opencvs():
    import cv2
    rng = np.random.default_rng(3)
    for trial in range(40):
        H, W = rng.integers(1, 48, 2)
        m = rng.random((H, W)) < rng.uniform(0.2, 0.7)
        n, lab, stats, _ = cv2.connectedComponentsWithStats(m.astype(np.uint8), 8, cv2.CV_32S)
        n2, lab2, area2 = P.plain_components8(m)
        assert n == n2 and np.array_equal(lab, lab2)
        assert np.array_equal(stats[:, cv2.CC_STAT_AREA], area2)


def test_fill_holes_is_scipys():
    import scipy.ndimage as ndi
    rng = np.random.default_rng(4)
    for trial in range(40):
        H, W = rng.integers(1, 40, 2)
        m = rng.random((H, W)) < rng.uniform(0.3, 0.8)
        assert np.array_equal(ndi.binary_fill_holes(m), P.plain_fill_holes(m))


def test_nearest_feature_ties_are_scipys():
    import scipy.ndimage as ndi
    rng = np.random.default_rng(5)
    for trial in range(30):
        H, W = rng.integers(1, 30, 2)
        cell = rng.random((H, W)) < rng.uniform(0.03, 0.5)
        if not cell.any():
            continue
        dist, ind = ndi.distance_transform_edt(~cell, return_indices=True)
        for lim_d in (2, 2.5, 4):
            lim = P.edge_dist_threshold(lim_d)
            for y in range(H):
                for x in range(W):
                    near = P.plain_nearest_cell(cell, y, x, lim)
                    if dist[y, x] < lim_d:
                        assert near == (ind[0, y, x], ind[1, y, x])
                    else:
                        assert near is None


def test_edge_dist_threshold():
    assert P.edge_dist_threshold(2) == 4        # d2 in {0..3}
    assert P.edge_dist_threshold(2.5) == 7      # sqrt(6) = 2.449 < 2.5 <= sqrt(7)
    assert P.edge_dist_threshold(0) == 0
