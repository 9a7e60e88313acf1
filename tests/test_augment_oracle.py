"""The augmentation oracle (oracle/augment_oracle.py) against the vectors made by the reference's own helpers
(tests/golden/make_augment_golden.py), and its first-principles restatements of cv2.warpAffine / SciPy's
map_coordinates, gaussian_filter and grey_dilation against those libraries."""
import os

import numpy as np
import pytest

from oracle import augment_oracle as A

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'augment.npz')


def augment_cases():
    g = np.load(GOLD)
    for name in g['names']:
        name = str(name)
        c = {k.split('/', 1)[1]: g[k] for k in g.files if k.startswith(name + '/')}
        yield name, c


@pytest.mark.parametrize('case', list(augment_cases()), ids=lambda c: c[0])
def test_oracle_matches_reference_vectors(case):
    name, c = case
    flip, rot = (int(c['flip_rot'][0]), int(c['flip_rot'][1])), int(c['flip_rot'][2])
    for t in range(c['img'].shape[0]):
        img, seg = A.augment_frame(c['img'][t], c['seg'][t], c['contrast'][t], c['brightness'][t], c.get('affine'),
                                   c.get('coords'), flip, rot)
        assert np.array_equal(img, c['out_img'][t]), (name, t)
        assert np.array_equal(seg, c['out_seg'][t]), (name, t)


@pytest.mark.parametrize('case', [c for c in augment_cases() if 'rand2' in c[1]], ids=lambda c: c[0])
def test_elastic_field_matches_reference_vectors(case):
    name, c = case
    H, W = c['img'].shape[1:]
    coords = A.elastic_coords(c['rand2'], W * 2, W * 0.15)
    np.testing.assert_allclose(coords, c['coords'], rtol=0, atol=1e-10)


def test_warp_affine_is_opencvs():
    import cv2
    rng = np.random.default_rng(0)
    for trial in range(12):
        H, W = rng.integers(8, 70, 2)
        img = (rng.random((H, W)) * 1000).astype(np.float32)
        c, s = np.float32([W, H]) // 2, min(H, W) // 3
        p1 = np.float32([c + s, [c[0] + s, c[1] - s], c - s])
        M = cv2.getAffineTransform(p1, p1 + rng.uniform(-W * 0.08, W * 0.08, p1.shape).astype(np.float32))
        assert np.array_equal(cv2.warpAffine(img, M, (W, H), borderMode=cv2.BORDER_REFLECT_101), A.warp_affine_linear(img, M))
        seg = rng.integers(0, 5, (H, W)).astype(np.float32)
        ref = cv2.warpAffine(seg, M, (W, H), borderMode=cv2.BORDER_CONSTANT, borderValue=-1, flags=cv2.INTER_NEAREST)
        assert np.array_equal(ref, A.warp_affine_nearest(seg, M))


def test_map_coordinates_and_filters_are_scipys():
    from scipy.ndimage import map_coordinates, gaussian_filter
    rng = np.random.default_rng(1)
    for trial in range(10):
        H, W = rng.integers(4, 50, 2)
        img = (rng.random((H, W)) * 1000).astype(np.float32)
        yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing='ij')
        cy, cx = yy + rng.normal(0, 5, (H, W)), xx + rng.normal(0, 5, (H, W))
        if trial % 3 == 0:
            cy, cx = cy * 3 - H, cx * 3 - W
        ref = map_coordinates(img, (cy.reshape(-1, 1), cx.reshape(-1, 1)), order=1, mode='reflect').reshape(H, W)
        assert np.array_equal(ref, A.map_linear_reflect(img, cy, cx))
        seg = rng.integers(0, 5, (H, W)).astype(np.float32)
        ref = map_coordinates(seg, (cy.reshape(-1, 1), cx.reshape(-1, 1)), order=0, mode='constant', cval=-1).reshape(H, W)
        assert np.array_equal(ref, A.map_nearest_constant(seg, cy, cx))
        f = rng.random((H, W)) * 2 - 1
        sigma = rng.uniform(0.5, 6)
        np.testing.assert_allclose(A.gaussian_filter_reflect(f, sigma), gaussian_filter(f, sigma), rtol=0, atol=1e-14)
