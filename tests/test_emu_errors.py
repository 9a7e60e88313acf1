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


def test_model_handle_argument_checks(lib):
    """the stateful entry points: use before binding, bad unroll lengths, bad state selectors, graph mode off in the host build"""
    from lstm_unet_b200 import _lib
    net = {'down_conv_kernels': [[(3, 4)], [(3, 6)]], 'lstm_kernels': [[(3, 3)], [(3, 5)]], 'up_conv_kernels': [[(3, 4)], [(3, 4), (1, 3)]]}
    cfg = _lib.make_config(net, 'NCHW', False, batch=2, max_t=2, height=8, width=8, precision='bf16x3', engine='simt', train=True)
    h = ctypes.c_void_p()
    assert lib.lu_create(None, ctypes.byref(h)) != 0
    bad = _lib.make_config(net, 'NCHW', False, batch=2, max_t=2, height=8, width=8, precision='bf16x3', engine='simt')
    bad.n_levels = 9
    assert lib.lu_create(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert lib.lu_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
    x = np.zeros((2, 2, 1, 8, 8), np.float32)
    out = np.zeros((2, 2, 3, 8, 8), np.float32)
    # nothing bound yet
    assert lib.lu_forward(h, x.ctypes.data, 2, 0, out.ctypes.data, out.ctypes.data, None) != 0
    assert b'bind' in lib.lu_last_error()
    nb = ctypes.c_size_t()
    assert lib.lu_workspace_bytes(h, ctypes.byref(nb)) == 0
    raw, ws = _aligned(nb.value, 1024)
    assert lib.lu_bind_workspace(h, ws, nb.value - 1, None) != 0 and b'too small' in lib.lu_last_error()
    assert lib.lu_bind_workspace(h, ws + 8, nb.value, None) != 0 and b'aligned' in lib.lu_last_error()
    assert lib.lu_bind_workspace(h, ws, nb.value, None) == 0
    nt, ne, ntr = ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64()
    assert lib.lu_param_count(h, ctypes.byref(nt), ctypes.byref(ne), ctypes.byref(ntr)) == 0
    params = np.zeros(ne.value, np.float32)
    assert lib.lu_bind_params(h, params.ctypes.data) == 0
    for T in (0, 3):                                                      # outside [1, max_t]
        assert lib.lu_forward(h, x.ctypes.data, T, 0, out.ctypes.data, out.ctypes.data, None) != 0
    assert lib.lu_forward(h, None, 2, 0, out.ctypes.data, out.ctypes.data, None) != 0
    # backward before any forward / after an inference forward
    lab, loss, grads = np.zeros((2, 2, 1, 8, 8), np.float32), np.zeros(1, np.float32), np.zeros(ntr.value, np.float32)
    cw = (ctypes.c_float * 3)(0.15, 0.25, 0.6)
    assert lib.lu_loss_backward(h, lab.ctypes.data, cw, loss.ctypes.data, grads.ctypes.data, None) != 0
    assert lib.lu_forward(h, x.ctypes.data, 2, 0, out.ctypes.data, out.ctypes.data, None) == 0
    assert lib.lu_loss_backward(h, lab.ctypes.data, cw, loss.ctypes.data, grads.ctypes.data, None) != 0
    assert b'training=1' in lib.lu_last_error()
    assert lib.lu_loss_backward(h, lab.ctypes.data, cw, loss.ctypes.data, None, None) == 0           # loss only is fine
    # state selectors
    shp = (ctypes.c_int64 * 4)()
    assert lib.lu_state_shape(h, 5, 0, shp) != 0 and lib.lu_state_shape(h, 0, 3, shp) != 0
    assert lib.lu_state_shape(h, 1, 0, shp) == 0 and tuple(shp) == (2, 5, 4, 4)
    eff = ctypes.c_int32(7)
    assert lib.lu_set_graph_mode(h, 1, ctypes.byref(eff)) == 0 and eff.value == 0                    # no graphs in the host build
    assert lib.lu_adam_step(h, grads.ctypes.data, grads.ctypes.data, grads.ctypes.data, 1e-3, 0.9, 0.999, 1e-7, 0, None) != 0
    assert lib.lu_destroy(h) == 0
