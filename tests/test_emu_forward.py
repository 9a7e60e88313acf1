"""Host-logic parity: the library's plan / tables / weight packing / epilogues, executed by the TEST-ONLY host build
(scalar mirror engine), against the oracle.  The same tables drive the tcgen05 kernel on the GPU (tests -m gpu)."""
import numpy as np
import pytest
import torch

from oracle import lstm_unet_oracle as O
from tests.emu_backend import emu_session, emu_forward

NET_A = {   # exercises channel padding (non multiples of 64), 4 levels
    'down_conv_kernels': [[(3, 8), (3, 8)], [(3, 12), (3, 12)], [(3, 12), (3, 12)], [(3, 16), (3, 16)]],
    'lstm_kernels': [[(5, 8)], [(5, 12)], [(5, 12)], [(5, 16)]],
    'up_conv_kernels': [[(3, 12), (3, 12)], [(3, 8), (3, 8)], [(3, 8), (3, 8)], [(3, 4), (3, 4), (1, 3)]],
}
NET_B = {   # 2 levels, 3x3 lstm, two lstm layers at level 0
    'down_conv_kernels': [[(3, 6)], [(3, 10), (3, 10)]],
    'lstm_kernels': [[(3, 5), (3, 7)], [(5, 9)]],
    'up_conv_kernels': [[(3, 6)], [(3, 5), (1, 3)]],
}


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def oracle_params_np(net, seed):
    p = O.init_params(net, seed=seed, randomize_bn=True)
    return p, {k: v.numpy() for k, v in p.items()}


@pytest.mark.parametrize("net,H,W,pad,a_mode", [
    (NET_A, 16, 24, False, 'halo'),
    (NET_A, 16, 24, False, 'direct'),
    (NET_A, 19, 21, True, 'halo'),
    (NET_B, 10, 12, False, 'halo'),
    (NET_B, 9, 7, True, 'direct'),
])
def test_forward_inference_bf16x3_matches_oracle(net, H, W, pad, a_mode):
    B, T = 2, 2
    p_t, p_np = oracle_params_np(net, 3)
    ora = O.OracleNet(net, 'NCHW', pad, params=p_t)
    sess = emu_session(net, data_format='NCHW', pad_image=pad, batch=B, max_t=T, height=H, width=W,
                       precision='bf16x3', a_mode=a_mode)
    sess.set_params(p_np)
    rng = np.random.default_rng(0)
    for call in range(2):          # second call exercises the stateful carry
        x = rng.standard_normal((B, T, 1, H, W)).astype(np.float32)
        ref_l, ref_s = ora(torch.from_numpy(x), False)
        got_l, got_s = emu_forward(sess, x, False)
        assert rel_err(got_l, ref_l.numpy()) < 1e-3, (call, rel_err(got_l, ref_l.numpy()))
        assert rel_err(got_s, ref_s.numpy()) < 1e-3
    sess.close()


def test_forward_training_bn_and_moving_stats():
    net, B, T, H, W = NET_A, 2, 2, 16, 16
    p_t, p_np = oracle_params_np(net, 5)
    ora = O.OracleNet(net, 'NCHW', False, params=p_t)
    sess = emu_session(net, data_format='NCHW', pad_image=False, batch=B, max_t=T, height=H, width=W,
                       precision='bf16x3', train=True)
    sess.set_params(p_np)
    x = np.random.default_rng(1).standard_normal((B, T, 1, H, W)).astype(np.float32)
    ref_l, _ = ora(torch.from_numpy(x), True)
    got_l, _ = emu_forward(sess, x, True)
    assert rel_err(got_l, ref_l.numpy()) < 1e-3, rel_err(got_l, ref_l.numpy())
    got_p = sess.get_params()
    for name in p_np:
        if 'moving' in name:
            np.testing.assert_allclose(got_p[name], ora.params[name].numpy(), rtol=2e-4, atol=2e-5, err_msg=name)
    sess.close()


def test_inference_after_training_forward_uses_the_new_moving_statistics():
    """model(x, True) moves the BatchNorm moving statistics without any optimizer step (the reference's unit_test loops,
    Networks.py:256-277); the next model(x, False) must normalise with them, not with the constants folded earlier."""
    net, B, T, H, W = NET_B, 2, 2, 10, 12
    p_t, p_np = oracle_params_np(net, 7)
    ora = O.OracleNet(net, 'NCHW', False, params=p_t)
    sess = emu_session(net, data_format='NCHW', pad_image=False, batch=B, max_t=T, height=H, width=W, precision='bf16x3')
    sess.set_params(p_np)
    rng = np.random.default_rng(2)
    for call, training in enumerate((False, True, True, False, False)):
        x = (3.0 * rng.standard_normal((B, T, 1, H, W)) + 1.0).astype(np.float32)     # far from the moving statistics
        ref_l, _ = ora(torch.from_numpy(x), training)
        got_l, _ = emu_forward(sess, x, training)
        assert rel_err(got_l, ref_l.numpy()) < 1e-3, (call, training, rel_err(got_l, ref_l.numpy()))
    sess.close()


def test_forward_bf16_mode_is_close():
    net, B, T, H, W = NET_A, 1, 2, 16, 16
    p_t, p_np = oracle_params_np(net, 7)
    ora = O.OracleNet(net, 'NCHW', False, params=p_t)
    sess = emu_session(net, data_format='NCHW', pad_image=False, batch=B, max_t=T, height=H, width=W, precision='bf16')
    sess.set_params(p_np)
    x = np.random.default_rng(2).standard_normal((B, T, 1, H, W)).astype(np.float32)
    ref_l, _ = ora(torch.from_numpy(x), False)
    got_l, _ = emu_forward(sess, x, False)
    assert rel_err(got_l, ref_l.numpy()) < 5e-2, rel_err(got_l, ref_l.numpy())   # bf16 operands: ~3 decimal digits
    sess.close()


def test_channels_last_softmax_quirk():
    net, B, T, H, W = NET_B, 2, 1, 8, 8
    p_t, p_np = oracle_params_np(net, 9)
    ora = O.OracleNet(net, 'NHWC', False, params=p_t)
    sess = emu_session(net, data_format='NHWC', pad_image=False, batch=B, max_t=T, height=H, width=W,
                       precision='bf16x3')
    sess.set_params(p_np)
    x = np.random.default_rng(3).standard_normal((B, T, H, W, 1)).astype(np.float32)
    ref_l, ref_s = ora(torch.from_numpy(x), False)
    got_l, got_s = emu_forward(sess, x, False)
    assert got_l.shape == (B, T, H, W, 3)
    assert rel_err(got_l, ref_l.numpy()) < 1e-3
    assert rel_err(got_s, ref_s.numpy()) < 1e-3          # softmax over the batch axis, as the reference computes it
    sess.close()


def test_state_mask_get_set():
    net, B, T, H, W = NET_B, 2, 2, 8, 8
    p_t, p_np = oracle_params_np(net, 11)
    ora = O.OracleNet(net, 'NCHW', False, params=p_t)
    sess = emu_session(net, data_format='NCHW', pad_image=False, batch=B, max_t=T, height=H, width=W,
                       precision='bf16x3')
    sess.set_params(p_np)
    x = np.random.default_rng(4).standard_normal((B, T, 1, H, W)).astype(np.float32)
    ora(torch.from_numpy(x), False)
    emu_forward(sess, x, False)
    mask = np.array([1.0, 0.0], dtype=np.float32)
    ora.reset_states_per_batch(mask)
    sess.reset_states(mask.ctypes.data)
    ref_states = ora.get_states()
    for (lvl, lay) in [(0, 0), (0, 1), (1, 0)]:
        shp = sess.state_shape(lvl, lay)
        for which in (0, 1):
            out = np.zeros(shp, dtype=np.float32)
            sess.get_state(lvl, lay, which, out.ctypes.data)
            ref = ref_states[lvl][lay][which]
            assert out.shape == ref.shape
            assert np.abs(out - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
            assert np.all(out[1] == 0)
    # set_state(NULL) zeroes; set_state(array) round-trips
    shp = sess.state_shape(1, 0)
    a = np.random.default_rng(5).standard_normal(shp).astype(np.float32)
    sess.set_state(1, 0, 1, a.ctypes.data)
    out = np.zeros(shp, dtype=np.float32)
    sess.get_state(1, 0, 1, out.ctypes.data)
    np.testing.assert_array_equal(out, a)
    sess.set_state(1, 0, 0, None)
    sess.get_state(1, 0, 0, out.ctypes.data)
    assert np.all(out == 0)
    sess.close()


def test_config_errors():
    from lstm_unet_b200 import _lib
    bad = dict(NET_A, lstm_kernels=NET_A['lstm_kernels'][:3])
    with pytest.raises(ValueError):
        _lib.make_config(bad)
    from lstm_unet_b200.session import LuError
    with pytest.raises(LuError):     # REFLECT pad needs pad < dim
        emu_session(NET_A, pad_image=True, batch=1, max_t=1, height=8, width=8)


def test_variable_unroll_length_and_minimal_sizes():
    """T may change between calls (T <= max_t): frames are addressed b*T+t per call; states carry over.  Also the
    smallest legal frame (8x8 for total_stride 8) and a 1x1-sample batch."""
    net = NET_A
    p_t, p_np = oracle_params_np(net, 13)
    ora = O.OracleNet(net, 'NCHW', False, params=p_t)
    sess = emu_session(net, data_format='NCHW', pad_image=False, batch=1, max_t=3, height=8, width=8, precision='bf16x3')
    sess.set_params(p_np)
    rng = np.random.default_rng(6)
    for T in (3, 1, 2):
        x = rng.standard_normal((1, T, 1, 8, 8)).astype(np.float32)
        ref_l, _ = ora(torch.from_numpy(x), False)
        got_l, _ = emu_forward(sess, x, False)
        assert got_l.shape == (1, T, 3, 8, 8)
        assert rel_err(got_l, ref_l.numpy()) < 1e-3, T
    from lstm_unet_b200.session import LuError
    with pytest.raises(LuError):
        emu_forward(sess, rng.standard_normal((1, 4, 1, 8, 8)).astype(np.float32), False)     # T > max_t
    sess.close()


def test_channels_last_training_step():
    """NHWC ('NWHC' in the reference's CLI, train2D.py:319): labels (B,T,H,W,1); logits gradients are layout-free."""
    net, B, T, H, W = NET_B, 2, 1, 8, 8
    p_t, p_np = oracle_params_np(net, 15)
    ora = O.OracleNet(net, 'NWHC', False, params=p_t)
    sess = emu_session(net, data_format='NWHC', pad_image=False, batch=B, max_t=T, height=H, width=W,
                       precision='bf16x3', train=True)
    sess.set_params(p_np)
    rng = np.random.default_rng(8)
    x = rng.standard_normal((B, T, H, W, 1)).astype(np.float32)
    lab = rng.integers(-1, 3, size=(B, T, H, W, 1)).astype(np.float32)
    names = ora.trainable_names()
    m = {n: torch.zeros_like(ora.params[n]) for n in names}
    v = {n: torch.zeros_like(ora.params[n]) for n in names}
    ref_loss, ref_logits, _, ref_grads = O.train_step(ora, torch.from_numpy(x), torch.from_numpy(lab), [0.15, 0.25, 0.6], m, v, 1, 0.0)
    logits, _ = emu_forward(sess, x, True)
    assert rel_err(logits, ref_logits.numpy()) < 1e-3
    grads = np.zeros(sess.n_trainable, np.float32)
    loss = np.zeros(1, np.float32)
    sess.loss_backward(lab.ctypes.data, [0.15, 0.25, 0.6], loss.ctypes.data, grads.ctypes.data)
    assert abs(float(loss[0]) - float(ref_loss)) < 1e-4
    e = [e for e in sess.layout if e['name'] == 'UpLayers/1/Conv/1/kernel'][0]
    g = grads[e['offset']:e['offset'] + e['count']].reshape(e['shape'])
    assert rel_err(g, ref_grads[e['name']].numpy()) < 5e-3
    sess.close()


def test_forward_fp16_mode_is_8x_closer_than_bf16_and_inference_only():
    """precision='fp16' (LU_PREC_FP16): fp16 operands through the same tables / epilogues.  Against the fp32 oracle the
    error must sit well below the bf16 mode's (11 vs 8 mantissa bits) and below the north_star's 1e-3 on this small
    network; creating a training handle in this mode is rejected."""
    from lstm_unet_b200.session import LuError
    net, B, T, H, W = NET_A, 2, 2, 19, 21
    p_t, p_np = oracle_params_np(net, 7)
    x = np.random.default_rng(2).standard_normal((B, T, 1, H, W)).astype(np.float32)
    errs = {}
    for prec in ('fp16', 'bf16'):
        ora = O.OracleNet(net, 'NCHW', True, params={k: v.clone() for k, v in p_t.items()})
        sess = emu_session(net, data_format='NCHW', pad_image=True, batch=B, max_t=T, height=H, width=W, precision=prec)
        sess.set_params(p_np)
        e = 0.0
        for call in range(2):
            ref_l, ref_s = ora(torch.from_numpy(x), False)
            got_l, got_s = emu_forward(sess, x, False)
            e = max(e, rel_err(got_l, ref_l.numpy()), rel_err(got_s, ref_s.numpy()))
        errs[prec] = e
        sess.close()
    assert errs['fp16'] < 1e-3, errs
    assert errs['fp16'] < errs['bf16'] / 3, errs
    with pytest.raises(LuError):
        emu_session(net, data_format='NCHW', pad_image=False, batch=1, max_t=1, height=16, width=16, precision='fp16', train=True)


def test_fp16_conversions_match_numpy_float16():
    """The portable fp32 <-> fp16 conversions of the host build (lu_f2half / lu_half2f: round to nearest even,
    subnormals, overflow) against numpy's float16, through set_state / get_state of an fp16 handle (h is stored in
    the operand format)."""
    net = {'down_conv_kernels': [[(3, 4)]], 'lstm_kernels': [[(3, 64)]], 'up_conv_kernels': [[(3, 4), (1, 3)]]}
    B, H, W, F = 1, 16, 16, 64
    sess = emu_session(net, data_format='NCHW', pad_image=False, batch=B, max_t=1, height=H, width=W, precision='fp16')
    rng = np.random.default_rng(0)
    v = rng.standard_normal(B * F * H * W).astype(np.float32)
    v[:4000] *= 10.0 ** rng.uniform(-9, 5, 4000).astype(np.float32)           # subnormals ... overflow
    special = np.array([0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e6, -1e6, 5.9604645e-8, 2.9802322e-8, 2.98023224e-8 * 1.0001,
                        6.1035156e-5, 6.0975552e-5, 1.0 + 2.0 ** -11, 1.0 + 2.0 ** -11 + 2.0 ** -20, 1.0 + 3 * 2.0 ** -11,
                        2047.5, 2048.5, 0.33325195], dtype=np.float32)
    v[4000:4000 + special.size] = special
    x = v.reshape(B, F, H, W)
    sess.set_state(0, 0, 0, x.ctypes.data)
    out = np.zeros_like(x)
    sess.get_state(0, 0, 0, out.ctypes.data)
    with np.errstate(over='ignore'):
        want = x.astype(np.float16).astype(np.float32)
    np.testing.assert_array_equal(out, want)
    sess.close()


@pytest.mark.parametrize("fmt,C,H,W,pad", [('NHWC', 3, 35, 35, True), ('NCHW', 2, 16, 24, False)])
def test_multi_channel_images(fmt, C, H, W, pad):
    """in_channels > 1 (the reference's unit_test feeds d=3 channels, channels-last, 35x35 with pad_image:
    Networks.py:266-270): the image is an ordinary NHWC source of the level-0 ConvLSTM and of the last decoder block's
    skip connection."""
    net, B, T = NET_B, 2, 2
    p_t = O.init_params(net, seed=11, randomize_bn=True, in_channels=C)
    ora = O.OracleNet(net, fmt, pad, params=p_t, in_channels=C)
    sess = emu_session(net, data_format=fmt, pad_image=pad, batch=B, max_t=T, height=H, width=W, precision='bf16x3',
                       in_channels=C)
    sess.set_params({k: v.numpy() for k, v in p_t.items()})
    rng = np.random.default_rng(0)
    shape = (B, T, C, H, W) if fmt == 'NCHW' else (B, T, H, W, C)
    for call in range(2):
        x = rng.standard_normal(shape).astype(np.float32)
        ref_l, ref_s = ora(torch.from_numpy(x), False)
        got_l, got_s = emu_forward(sess, x, False)
        assert got_l.shape == tuple(ref_l.shape)
        assert rel_err(got_l, ref_l.numpy()) < 1e-3 and rel_err(got_s, ref_s.numpy()) < 1e-3
    sess.close()
