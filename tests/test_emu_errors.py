"""Error behaviour of the stateless C entry points (post-processing, SEG measure, augmentation) through the TEST-ONLY host
build: every misuse returns non-zero with a message in lu_last_error() instead of touching memory."""
import ctypes

import numpy as np
import pytest

from tests.emu_backend import build_emu


@pytest.fixture(scope='module')
def lib():
    from lstm_unet_b200 import _lib
    return _lib.load_library(build_emu())


def _aligned(nbytes, align=256):
    raw = np.zeros(nbytes + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw, raw.ctypes.data + off


def test_postprocess_argument_checks(lib):
    from lstm_unet_b200 import _lib
    nb = ctypes.c_size_t()
    assert lib.lu_post_workspace_bytes(0, 8, 8, ctypes.byref(nb)) != 0 and b'frames' in lib.lu_last_error()
    assert lib.lu_post_workspace_bytes(1, 8, 8, ctypes.byref(nb)) == 0 and nb.value > 0
    sm = np.zeros((3, 8, 8), np.float32)
    out = np.zeros((8, 8), np.uint16)
    pp = _lib.lu_post_params(0.2, 4, 1, 100, 0, 1)
    raw, ws = _aligned(nb.value)
    assert lib.lu_postprocess(None, 1, 8, 8, ctypes.byref(pp), out.ctypes.data, None, ws, nb.value, None) != 0
    assert lib.lu_postprocess(sm.ctypes.data, 1, 8, 8, ctypes.byref(pp), out.ctypes.data, None, ws, nb.value - 1, None) != 0
    assert b'workspace too small' in lib.lu_last_error()
    assert lib.lu_postprocess(sm.ctypes.data, 1, 8, 8, ctypes.byref(pp), out.ctypes.data, None, ws + 8, nb.value, None) != 0
    assert b'aligned' in lib.lu_last_error()
    bad = _lib.lu_post_params(0.2, 4, 1, 100, 8, 1)                      # FOV >= width
    assert lib.lu_postprocess(sm.ctypes.data, 1, 8, 8, ctypes.byref(bad), out.ctypes.data, None, ws, nb.value, None) != 0
    bad = _lib.lu_post_params(0.2, 10 ** 6, 1, 100, 0, 1)                # absurd edge distance
    assert lib.lu_postprocess(sm.ctypes.data, 1, 8, 8, ctypes.byref(bad), out.ctypes.data, None, ws, nb.value, None) != 0
    assert lib.lu_postprocess(sm.ctypes.data, 1, 8, 8, ctypes.byref(pp), out.ctypes.data, None, ws, nb.value, None) == 0
    assert not out.any()                                                  # all-zero soft-max: argmax is class 0


def test_seg_and_augment_argument_checks(lib):
    from lstm_unet_b200 import _lib
    nb = ctypes.c_size_t()
    assert lib.lu_seg_workspace_bytes(1, 0, 4, ctypes.byref(nb)) != 0
    assert lib.lu_seg_workspace_bytes(1, 4, 4, ctypes.byref(nb)) == 0
    lab, lg, res = np.zeros((1, 1, 4, 4), np.float32), np.zeros((1, 3, 4, 4), np.float32), np.zeros(4, np.float64)
    raw, ws = _aligned(nb.value)
    assert lib.lu_seg_measure(lab.ctypes.data, lg.ctypes.data, 1, 4, 4, 1, None, ws, nb.value, None) != 0
    assert lib.lu_seg_measure(lab.ctypes.data, lg.ctypes.data, 1, 4, 4, 1, res.ctypes.data, ws, 16, None) != 0
    assert lib.lu_seg_measure(lab.ctypes.data, lg.ctypes.data, 1, 4, 4, 1, res.ctypes.data, ws, nb.value, None) == 0
    assert res[1] == 0 and res[2] == 16 and res[3] == 16                  # no truth objects; class 0 everywhere is correct
    ap = _lib.lu_aug_params()
    ap.frames, ap.H, ap.W, ap.randomize, ap.elastic, ap.rot90 = 1, 4, 4, 1, 0, 0
    assert lib.lu_aug_workspace_bytes(1, 4, 4, ctypes.byref(nb)) == 0
    raw2, ws2 = _aligned(nb.value)
    img, seg = np.zeros((1, 4, 4), np.float32), np.zeros((1, 4, 4), np.float32)
    oi, os_ = np.zeros_like(img), np.zeros_like(seg)
    args = (img.ctypes.data, seg.ctypes.data)
    # randomize without the contrast / brightness arrays, elastic without coordinates, bad rotation
    assert lib.lu_augment_sequence(*args, None, None, None, ctypes.byref(ap), oi.ctypes.data, os_.ctypes.data, ws2, nb.value, None) != 0
    ap.randomize, ap.elastic = 0, 1
    assert lib.lu_augment_sequence(*args, None, None, None, ctypes.byref(ap), oi.ctypes.data, os_.ctypes.data, ws2, nb.value, None) != 0
    ap.elastic, ap.rot90 = 0, 7
    assert lib.lu_augment_sequence(*args, None, None, None, ctypes.byref(ap), oi.ctypes.data, os_.ctypes.data, ws2, nb.value, None) != 0
    ap.rot90 = 0
    assert lib.lu_augment_sequence(*args, None, None, None, ctypes.byref(ap), oi.ctypes.data, os_.ctypes.data, ws2, nb.value, None) == 0
    assert lib.lu_elastic_coords(None, None, 1, 4, 4, 1.0, None, None, None) != 0
