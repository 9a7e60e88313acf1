"""Pins the torch-CPU oracle against the independent numpy statement of the operator semantics and against
the shape contracts the reference's own unit_test methods print (Networks.py:100-119,155-175,256-277)."""
import numpy as np
import pytest
import torch

from oracle import lstm_unet_oracle as O
from oracle import np_semantics as S

SMALL = {
    'down_conv_kernels': [[(3, 8), (3, 8)], [(3, 12), (3, 12)], [(3, 12), (3, 12)], [(3, 16), (3, 16)]],
    'lstm_kernels': [[(5, 8)], [(5, 12)], [(5, 12)], [(5, 16)]],
    'up_conv_kernels': [[(3, 12), (3, 12)], [(3, 8), (3, 8)], [(3, 8), (3, 8)], [(3, 4), (3, 4), (1, 3)]],
}


def nchw(a):
    return torch.from_numpy(np.ascontiguousarray(np.transpose(a, (0, 3, 1, 2))))


def nhwc(t):
    return t.permute(0, 2, 3, 1).numpy()


@pytest.mark.parametrize("k,stride,H,W", [(3, 1, 6, 7), (3, 2, 8, 6), (5, 1, 6, 5), (1, 1, 4, 4), (3, 2, 7, 5)])
def test_conv_same_matches_numpy(k, stride, H, W):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, H, W, 3))
    w = rng.standard_normal((k, k, 3, 4))
    b = rng.standard_normal(4)
    ref = S.conv2d_same_nhwc(x, w, b, stride)
    got = nhwc(O.conv2d_same(nchw(x), torch.from_numpy(w), torch.from_numpy(b), stride))
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)


def test_stride2_same_is_asymmetric():
    # k=3, s=2, even input pads (0 before, 1 after): out[j] reads in[2j..2j+2]
    assert O.tf_same_pad(8, 3, 2) == (0, 1)
    assert O.tf_same_pad(8, 3, 1) == (1, 1)
    assert O.tf_same_pad(8, 5, 1) == (2, 2)


def test_bilinear_matches_numpy():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 5, 4, 3))
    ref = S.bilinear_up_nhwc(x, 2)
    got = nhwc(O.resize_bilinear(nchw(x), 2))
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
    # closed form for f=2 (App. A.5)
    np.testing.assert_allclose(ref[:, 2, :, :], S.bilinear_up_nhwc(x, 2)[:, 2], rtol=0)
    row = 0.25 * x[:, 0] + 0.75 * x[:, 1]
    np.testing.assert_allclose(S.bilinear_up_nhwc(x[:, :, :1], 2)[:, 2, 0], row[:, 0], rtol=1e-12)
    assert O.resize_bilinear(nchw(x), 1) is not None and O.resize_bilinear(nchw(x), 1).shape == nchw(x).shape


def test_convlstm_sequence_matches_numpy():
    rng = np.random.default_rng(2)
    B, T, H, W, Ci, Fo, k = 2, 3, 6, 5, 2, 3, 5
    x = rng.standard_normal((B, T, H, W, Ci))
    wk = rng.standard_normal((k, k, Ci, 4 * Fo)) * 0.2
    wr = rng.standard_normal((k, k, Fo, 4 * Fo)) * 0.2
    b = rng.standard_normal(4 * Fo) * 0.5
    h = np.zeros((B, H, W, Fo))
    c = np.zeros((B, H, W, Fo))
    ref = []
    for t in range(T):
        h, c = S.convlstm_step_nhwc(x[:, t], h, c, wk, wr, b)
        ref.append(h)
    ref = np.stack(ref, 1)
    net_params = {'down_conv_kernels': [[(3, 2)]], 'lstm_kernels': [[(k, Fo)]], 'up_conv_kernels': [[(1, 3)]]}
    net = O.OracleNet(net_params, 'NCHW', False, dtype=torch.float64, in_channels=Ci)
    net.params['DownLayers/0/ConvLSTM/0/kernel'] = torch.from_numpy(wk)
    net.params['DownLayers/0/ConvLSTM/0/recurrent_kernel'] = torch.from_numpy(wr)
    net.params['DownLayers/0/ConvLSTM/0/bias'] = torch.from_numpy(b)
    x5 = torch.from_numpy(np.ascontiguousarray(np.transpose(x, (0, 1, 4, 2, 3))))
    got = net._conv_lstm(x5, 0, 0).permute(0, 1, 3, 4, 2).numpy()
    np.testing.assert_allclose(got, ref, rtol=1e-10, atol=1e-12)
    # stateful: final (h, c) stored, used by the next call
    np.testing.assert_allclose(net.states[0][0][0].permute(0, 2, 3, 1).numpy(), h, rtol=1e-10, atol=1e-12)
    h2, c2 = S.convlstm_step_nhwc(x[:, 0], h, c, wk, wr, b)
    got2 = net._conv_lstm(x5[:, :1], 0, 0).permute(0, 1, 3, 4, 2).numpy()
    np.testing.assert_allclose(got2[:, 0], h2, rtol=1e-10, atol=1e-12)


def test_bn_lrelu_reflect_ce_match_numpy():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 4, 5, 6)) * 2 + 1
    g, b = rng.standard_normal(6), rng.standard_normal(6)
    ref, mean, uvar = S.batchnorm_train_nhwc(x, g, b)
    mm, mv = torch.zeros(6, dtype=torch.float64), torch.ones(6, dtype=torch.float64)
    got = O.batchnorm(nchw(x), torch.from_numpy(g), torch.from_numpy(b), mm, mv, True)
    np.testing.assert_allclose(nhwc(got), ref, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(mm.numpy(), 0.01 * mean, rtol=1e-10)
    np.testing.assert_allclose(mv.numpy(), 0.99 + 0.01 * uvar, rtol=1e-10)
    np.testing.assert_allclose(O.leaky_relu(torch.from_numpy(x)).numpy(), S.leaky_relu(x))
    img = rng.standard_normal((2, 9, 7))
    ref_p = S.reflect_pad_hw(img, 3, 5, 2, 4)
    got_p = torch.nn.functional.pad(torch.from_numpy(img)[None], (2, 4, 3, 5), mode='reflect')[0].numpy()
    np.testing.assert_allclose(got_p, ref_p)
    labels = rng.integers(-1, 3, size=(2, 2, 1, 4, 5)).astype(np.float64)
    logits = rng.standard_normal((2, 2, 3, 4, 5))
    cw = [0.15, 0.25, 0.6]
    ref_l = S.weighted_ce(labels[:, :, 0], np.transpose(logits, (0, 1, 3, 4, 2)), cw)
    got_l = O.weighted_ce_loss(torch.from_numpy(labels), torch.from_numpy(logits), cw, True)
    np.testing.assert_allclose(float(got_l), ref_l, rtol=1e-12)


def test_reference_unit_test_shape_contract():
    # Networks.py:256-277: NHWC, pad_image=True, h=w=35, B=2, T=2, 4 stateful calls -> logits (2,2,35,35,3)
    net = O.OracleNet(SMALL, 'NHWC', True, in_channels=3)
    for _ in range(2):
        x = np.random.randn(2, 2, 35, 35, 3).astype(np.float32)
        logits, softmax = net(x, True)
        assert tuple(logits.shape) == (2, 2, 35, 35, 3)
        assert tuple(softmax.shape) == (2, 2, 35, 35, 3)
    # NHWC quirk: softmax normalises over the batch axis
    np.testing.assert_allclose(softmax.sum(0).numpy(), 1.0, rtol=1e-5)
    net = O.OracleNet(SMALL, 'NCHW', True)
    logits, softmax = net(np.random.randn(1, 2, 1, 35, 35).astype(np.float32), False)
    assert tuple(logits.shape) == (1, 2, 3, 35, 35)
    np.testing.assert_allclose(softmax.sum(2).numpy(), 1.0, rtol=1e-5)


def test_param_count_matches_survey():
    specs = O.build_param_specs(O.CTC_NET_PARAMS)
    tot = sum(int(np.prod(s)) for _, s, _ in specs)
    train = sum(int(np.prod(s)) for _, s, k in specs if k in O.TRAINABLE_KINDS)
    assert tot == 74613059 and train == 74606531
    assert len([1 for _, _, k in specs if k in O.TRAINABLE_KINDS]) == 78


def test_level_mismatch_raises():
    bad = dict(SMALL, lstm_kernels=SMALL['lstm_kernels'][:3])
    with pytest.raises(ValueError):
        O.OracleNet(bad)


def test_state_mask_and_swap():
    net = O.OracleNet(SMALL, 'NCHW', False, seed=1)
    x = torch.randn(2, 2, 1, 16, 16)
    net(x, False)
    st = net.get_states()
    assert st[0][0][0].shape == (2, 8, 16, 16)
    net.reset_states_per_batch(np.array([1.0, 0.0], dtype=np.float32))
    st2 = net.get_states()
    np.testing.assert_array_equal(st2[0][0][0][0], st[0][0][0][0])
    assert np.all(st2[0][0][1][1] == 0)
    a, _ = net(x, False)
    net.set_states(st2)
    b, _ = net(x, False)
    np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-6, atol=1e-6)


def test_train_step_decreases_loss_and_adam_formula():
    torch.manual_seed(0)
    net = O.OracleNet(SMALL, 'NCHW', False, dtype=torch.float64, seed=2)
    x = torch.randn(2, 2, 1, 16, 16, dtype=torch.float64)
    lab = torch.randint(-1, 3, (2, 2, 1, 16, 16)).double()
    names = net.trainable_names()
    m = {n: torch.zeros_like(net.params[n]) for n in names}
    v = {n: torch.zeros_like(net.params[n]) for n in names}
    p0 = net.params[names[0]].clone()
    loss, _, _, grads = O.train_step(net, x, lab, [0.15, 0.25, 0.6], m, v, 1, lr=1e-3)
    g = grads[names[0]]
    # first Adam step: m=(1-b1)g, v=(1-b2)g^2, lr_t = lr*sqrt(1-b2)/(1-b1)
    exp = p0 - 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9) * (0.1 * g) / (torch.sqrt(0.001 * g * g) + 1e-7)
    np.testing.assert_allclose(net.params[names[0]].numpy(), exp.numpy(), rtol=1e-9, atol=1e-12)
    assert float(loss) > 0


def test_bilinear_x2_agrees_with_opencv():
    """a third, independent implementation of the half-pixel-centre bilinear resize (tf.image.resize's TF2 default,
    SURVEY App. A.5): OpenCV's INTER_LINEAR uses the same sampling convention for float images"""
    cv2 = pytest.importorskip('cv2')
    import torch
    from oracle import lstm_unet_oracle as O
    rng = np.random.default_rng(0)
    for h, w in ((7, 9), (1, 5), (16, 3)):
        x = rng.standard_normal((1, 1, h, w)).astype(np.float32)
        mine = O.resize_bilinear(torch.from_numpy(x), 2)[0, 0].numpy()
        ref = cv2.resize(x[0, 0], (2 * w, 2 * h), interpolation=cv2.INTER_LINEAR)
        assert np.abs(mine - ref).max() < 1e-6


def test_blocks_on_their_own_chain_to_the_network():
    """oracle/blocks_oracle.py (the reference's blocks called alone, Networks.py:100-119,155-175) chained the way
    ULSTMnet2D.call chains them (Networks.py:233-245) reproduces OracleNet: one arithmetic on both sides."""
    from oracle import blocks_oracle as BO
    net = {'down_conv_kernels': [[(3, 8), (3, 8)], [(3, 12)]], 'lstm_kernels': [[(3, 6)], [(5, 10), (3, 10)]],
           'up_conv_kernels': [[(3, 8)], [(3, 6), (1, 3)]]}
    params = O.init_params(net, seed=4, randomize_bn=True, in_channels=2)
    ora = O.OracleNet(net, 'NCHW', False, params=params, in_channels=2)
    x = torch.randn(2, 3, 2, 16, 24, generator=torch.Generator().manual_seed(0))

    def sub(prefix):
        return {k.replace(prefix, prefix[:-2] + '0/'): v.clone() for k, v in params.items() if k.startswith(prefix)}
    d0 = BO.OracleDownBlock(net['down_conv_kernels'][0], net['lstm_kernels'][0], 2, 'NCHW', params=sub('DownLayers/0/'))
    d1 = BO.OracleDownBlock(net['down_conv_kernels'][1], net['lstm_kernels'][1], 1, 'NCHW', params=sub('DownLayers/1/'))
    u0 = BO.OracleUpBlock(net['up_conv_kernels'][0], 1, 'NCHW', False, params=sub('UpLayers/0/'))
    u1 = BO.OracleUpBlock(net['up_conv_kernels'][1], 2, 'NCHW', True, params=sub('UpLayers/1/'))
    for call in range(2):                                        # second call: stateful carry in both
        ref_logits, _ = ora(x + call, training=False)
        xi = x + call
        down0, skip0 = d0(xi, False)
        down1, skip1 = d1(down0, False)
        up = u0((skip1, skip0), False)
        logits = u1((up, xi.reshape(6, 2, 16, 24)), False)
        assert torch.allclose(logits.reshape(2, 3, 3, 16, 24), ref_logits, atol=1e-5), call
