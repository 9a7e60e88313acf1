"""Parity tests proper: the sm_100a tcgen05 path, called through the public Networks.ULSTMnet2D API (ctypes ->
C-ABI), against the CPU oracle on the same seeded inputs.  Tolerances: 1e-3 relative (max-abs error over max-abs
reference) for the bf16x3 parity mode -- the north_star's tolerance -- and 5e-2 for the plain-bf16 throughput mode
(bf16 operands carry 8 significant bits)."""
import numpy as np
import pytest
import torch

from oracle import lstm_unet_oracle as O

pytestmark = pytest.mark.gpu

NET_ODD = {
    'down_conv_kernels': [[(3, 8), (3, 8)], [(3, 12), (3, 12)], [(3, 12), (3, 12)], [(3, 16), (3, 16)]],
    'lstm_kernels': [[(5, 8)], [(5, 12)], [(5, 12)], [(5, 16)]],
    'up_conv_kernels': [[(3, 12), (3, 12)], [(3, 8), (3, 8)], [(3, 8), (3, 8)], [(3, 4), (3, 4), (1, 3)]],
}
NET_WIDE = {   # CTC channel structure at 1/4 width: every buffer a multiple of 64 except the 32-wide tail
    'down_conv_kernels': [[(3, 64), (3, 64)], [(3, 128), (3, 128)], [(3, 128), (3, 128)], [(3, 192), (3, 192)]],
    'lstm_kernels': [[(5, 64)], [(5, 128)], [(5, 128)], [(5, 192)]],
    'up_conv_kernels': [[(3, 128), (3, 128)], [(3, 64), (3, 64)], [(3, 32), (3, 32)], [(3, 16), (3, 16), (1, 3)]],
}
NET_TWO = {
    'down_conv_kernels': [[(3, 6)], [(3, 10), (3, 10)]],
    'lstm_kernels': [[(3, 5), (3, 7)], [(5, 9)]],
    'up_conv_kernels': [[(3, 6)], [(3, 5), (1, 3)]],
}


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def make_pair(net, data_format, pad, seed, **kw):
    from lstm_unet_b200.Networks import ULSTMnet2D
    params = O.init_params(net, seed=seed, randomize_bn=True)
    ora = O.OracleNet(net, data_format, pad, params=params)
    model = ULSTMnet2D(net, data_format, pad, **kw)
    model.set_weights_dict({k: v.numpy().copy() for k, v in params.items()})
    return ora, model


def test_library_is_cuda_build_and_loaded():
    from lstm_unet_b200 import _lib
    lib = _lib.load_library()
    assert lib.lu_is_cuda_build() == 1


@pytest.mark.parametrize("net,B,T,H,W,pad,a_mode", [
    (NET_ODD, 2, 2, 40, 48, True, 'halo'),
    (NET_ODD, 2, 3, 35, 35, True, 'halo'),       # the reference unit_test's 35x35 pad 8/13 case (Networks.py:266)
    (NET_ODD, 1, 2, 32, 24, False, 'direct'),
    (NET_TWO, 3, 2, 18, 22, False, 'halo'),
    (NET_WIDE, 1, 2, 64, 48, False, 'halo'),
])
def test_tcgen05_bf16x3_parity(net, B, T, H, W, pad, a_mode):
    ora, model = make_pair(net, 'NCHW', pad, 3, precision='bf16x3', a_mode=a_mode)
    rng = np.random.default_rng(0)
    for call in range(2):                          # second call: stateful carry of h, c
        x = rng.standard_normal((B, T, 1, H, W)).astype(np.float32)
        ref_l, ref_s = ora(torch.from_numpy(x), False)
        logits, softmax = model(x, training=False)
        assert tuple(logits.shape) == (B, T, 3, H, W)
        e1, e2 = rel_err(logits.numpy(), ref_l.numpy()), rel_err(softmax.numpy(), ref_s.numpy())
        assert e1 < 1e-3 and e2 < 1e-3, (call, e1, e2)
    assert model.launch_count() > 0


def test_inference_after_training_forwards_reference_unit_test_loop():
    """The reference's ULSTMnet2D.unit_test calls model(x, training=True) four times in a row (Networks.py:256-277): no
    optimizer step in between, the BatchNorm moving statistics move with every call, and an inference call afterwards
    normalises with the moved statistics."""
    B, T, H, W = 2, 2, 35, 35
    ora, model = make_pair(NET_ODD, 'NCHW', True, 9, precision='bf16x3')
    rng = np.random.default_rng(4)
    for call, training in enumerate((False, True, True, False, True, False)):
        x = (2.0 * rng.standard_normal((B, T, 1, H, W)) + 0.5).astype(np.float32)
        ref_l, _ = ora(torch.from_numpy(x), training)
        logits, _ = model(x, training=training)
        assert rel_err(logits.numpy(), ref_l.numpy()) < 1e-3, (call, training, rel_err(logits.numpy(), ref_l.numpy()))


def test_tcgen05_bf16_mode_close():
    ora, model = make_pair(NET_WIDE, 'NCHW', True, 5, precision='bf16')
    x = np.random.default_rng(1).standard_normal((2, 2, 1, 40, 56)).astype(np.float32)
    ref_l, _ = ora(torch.from_numpy(x), False)
    logits, _ = model(x, training=False)
    assert rel_err(logits.numpy(), ref_l.numpy()) < 5e-2


def test_training_mode_batchnorm_parity():
    ora, model = make_pair(NET_ODD, 'NCHW', False, 7, precision='bf16x3', train=True)
    x = np.random.default_rng(2).standard_normal((2, 2, 1, 32, 32)).astype(np.float32)
    ref_l, _ = ora(torch.from_numpy(x), True)
    logits, _ = model(x, training=True)
    assert rel_err(logits.numpy(), ref_l.numpy()) < 1e-3
    got = model.get_weights_dict()
    for name, ref in ora.params.items():
        if 'moving' in name:
            np.testing.assert_allclose(got[name], ref.numpy(), rtol=2e-4, atol=2e-5, err_msg=name)


def test_states_mask_get_set_and_streaming():
    ora, model = make_pair(NET_TWO, 'NCHW', False, 9, precision='bf16x3')
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 2, 1, 16, 16)).astype(np.float32)
    assert model.get_states()[0][0][0] is None     # before the first call (keras: states are None)
    ora(torch.from_numpy(x), False)
    model(x, False)
    mask = np.array([1.0, 0.0], dtype=np.float32)  # 1 = keep, 0 = reset (DataHandeling.py:378)
    ora.reset_states_per_batch(mask)
    model.reset_states_per_batch(mask)
    rs, gs = ora.get_states(), model.get_states()
    for lvl in range(2):
        for lay in range(len(rs[lvl])):
            for which in (0, 1):
                assert gs[lvl][lay][which].shape == rs[lvl][lay][which].shape
                assert np.abs(gs[lvl][lay][which] - rs[lvl][lay][which]).max() < 1e-4
                assert np.all(gs[lvl][lay][which][1] == 0)
    # streaming T=1 calls (Inference2D.py:45-59) equal one T=2 call from the same state
    saved = model.get_states()
    x2 = rng.standard_normal((2, 2, 1, 16, 16)).astype(np.float32)
    full, _ = model(x2, False)
    model.set_states(saved)
    a, _ = model(x2[:, :1], False)
    b, _ = model(x2[:, 1:], False)
    # (get/set_states round-trips h through fp32 and re-splits it into bf16 hi+lo: last-bit differences)
    np.testing.assert_allclose(np.concatenate([a.numpy(), b.numpy()], 1), full.numpy(), rtol=1e-4, atol=1e-5)
    with pytest.raises(ValueError):
        model(np.zeros((2, 1, 1, 24, 16), np.float32), False)      # B,H,W are frozen by the first call


def test_channels_last_quirk_and_errors():
    ora, model = make_pair(NET_TWO, 'NHWC', False, 11, precision='bf16x3')
    x = np.random.default_rng(4).standard_normal((2, 1, 16, 16, 1)).astype(np.float32)
    ref_l, ref_s = ora(torch.from_numpy(x), False)
    logits, softmax = model(x, False)
    assert rel_err(logits.numpy(), ref_l.numpy()) < 1e-3
    assert rel_err(softmax.numpy(), ref_s.numpy()) < 1e-3
    from lstm_unet_b200.Networks import ULSTMnet2D
    with pytest.raises(ValueError):
        ULSTMnet2D(dict(NET_TWO, lstm_kernels=NET_TWO['lstm_kernels'][:1]))


def test_linearity_of_logits_conv_at_scale():
    """Size-independent property at a larger shape than the oracle is run on: two runs from reset states with the
    same input are bit-identical (determinism), and halo vs direct staging agree to fp32 summation order."""
    from lstm_unet_b200.Networks import ULSTMnet2D
    params = O.init_params(NET_WIDE, seed=13, randomize_bn=True)
    w = {k: v.numpy() for k, v in params.items()}
    x = np.random.default_rng(5).standard_normal((2, 2, 1, 128, 160)).astype(np.float32)
    outs = []
    for a_mode in ('halo', 'halo', 'direct'):
        m = ULSTMnet2D(NET_WIDE, 'NCHW', True, precision='bf16x3', a_mode=a_mode)
        m.set_weights_dict(w)
        outs.append(m(x, False)[0].numpy())
    np.testing.assert_array_equal(outs[0], outs[1])
    assert rel_err(outs[2], outs[0]) < 1e-4


def test_unroll_length_may_change_and_grow():
    """T <= max_t reuses the handle; a longer unroll rebuilds it and carries weights and recurrent states over
    (Keras only freezes B, H, W of a stateful layer)."""
    ora, model = make_pair(NET_TWO, 'NCHW', False, 17, precision='bf16x3')
    rng = np.random.default_rng(9)
    for T in (2, 1, 4, 3):
        x = rng.standard_normal((2, T, 1, 16, 24)).astype(np.float32)
        ref_l, _ = ora(torch.from_numpy(x), False)
        logits, _ = model(x, False)
        assert tuple(logits.shape) == (2, T, 3, 16, 24)
        assert rel_err(logits.numpy(), ref_l.numpy()) < 1e-3, T
    model.close()


@pytest.mark.parametrize("precision", ['bf16', 'bf16x3'])
def test_cuda_graph_replay_is_bitwise_the_plain_forward(precision):
    """Inference2D's frame loop (B=1, T=1, stateful) replayed as CUDA graphs (lu_set_graph_mode): same bits as the
    plain launches over a sequence that alternates the recurrent-state ping-pong parity, with a state reset and a
    set_states in between, and with T=2 calls mixed in."""
    from lstm_unet_b200.Networks import ULSTMnet2D
    params = O.init_params(NET_WIDE, seed=5, randomize_bn=True)
    rng = np.random.default_rng(7)
    frames = [rng.standard_normal((1, 1, 1, 40, 56)).astype(np.float32) for _ in range(7)]
    pair = rng.standard_normal((1, 2, 1, 40, 56)).astype(np.float32)
    outs = {}
    for mode in (False, True):
        m = ULSTMnet2D(NET_WIDE, 'NCHW', True, precision=precision, cuda_graph=mode)
        m.set_weights_dict({k: v.numpy().copy() for k, v in params.items()})
        got = []
        m(pair, False)                                   # builds the session with max_t = 2
        assert m.graph_active == mode
        for i, f in enumerate(frames):
            lg, sm = m(f, False)
            got.append((lg.numpy().copy(), sm.numpy().copy()))
            if i == 2:
                m.reset_states_per_batch(np.zeros(1, np.float32))
            if i == 4:
                st = m.get_states()
                m.set_states(st)
                lg2, _ = m(pair, False)
                got.append((lg2.numpy().copy(),))
        n0 = m.launch_count(reset=True)
        m(frames[0], False)
        assert m.launch_count() > 0 and n0 > 0
        outs[mode] = got
        m.close()
    for a, b in zip(outs[False], outs[True]):
        for u, v in zip(a, b):
            assert np.array_equal(u, v)
    auto = ULSTMnet2D(NET_WIDE, 'NCHW', True, precision=precision)
    auto(frames[0], False)
    assert auto.graph_active                            # B*T = 1: on by default
    big = ULSTMnet2D(NET_WIDE, 'NCHW', True, precision=precision)
    big(np.zeros((2, 2, 1, 40, 56), np.float32), False)
    assert not big.graph_active


def test_fp16_operand_mode_meets_the_north_star_tolerance():
    """precision='fp16': fp16 operands at the bf16 tensor-core rate; 11-bit mantissa -> 1e-3 without the 3x split."""
    ora, model = make_pair(NET_WIDE, 'NCHW', True, 5, precision='fp16')
    rng = np.random.default_rng(1)
    for call in range(2):
        x = rng.standard_normal((2, 2, 1, 40, 56)).astype(np.float32)
        ref_l, ref_s = ora(torch.from_numpy(x), False)
        logits, softmax = model(x, training=False)
        assert rel_err(logits.numpy(), ref_l.numpy()) < 1e-3 and rel_err(softmax.numpy(), ref_s.numpy()) < 1e-3
    # h comes back through get_states in fp16
    rs, gs = ora.get_states(), model.get_states()
    assert rel_err(gs[1][0][0], rs[1][0][0]) < 1e-3 and rel_err(gs[1][0][1], rs[1][0][1]) < 1e-3
    from lstm_unet_b200.Networks import ULSTMnet2D
    m = ULSTMnet2D(NET_TWO, 'NCHW', False, precision='fp16', train=True)
    with pytest.raises(ValueError):
        m(np.zeros((1, 1, 1, 16, 16), np.float32), True)


def test_predict_batches_is_bitwise_the_successive_calls_and_results_are_caller_owned():
    """ULSTMnet2D.predict_batches (copies overlapped with compute on side streams) == model(x)[1].numpy() per batch,
    state carry included; arrays handed out by .numpy() are never overwritten by later read-backs."""
    from lstm_unet_b200.Networks import ULSTMnet2D
    params = O.init_params(NET_WIDE, seed=5, randomize_bn=True)
    w = {k: v.numpy().copy() for k, v in params.items()}
    rng = np.random.default_rng(3)
    batches = [rng.standard_normal((2, 2, 1, 40, 56)).astype(np.float32) for _ in range(6)]
    a = ULSTMnet2D(NET_WIDE, 'NCHW', True)
    a.set_weights_dict(w)
    seq = [a(x, False)[1].numpy() for x in batches]                 # kept in a list: must stay intact
    copies = [s.copy() for s in seq]
    b = ULSTMnet2D(NET_WIDE, 'NCHW', True)
    b.set_weights_dict(w)
    piped = list(b.predict_batches(iter(batches)))
    assert len(piped) == len(batches)
    for s, c, p in zip(seq, copies, piped):
        assert np.array_equal(s, c)
        assert np.array_equal(p, c)
    assert len({s.ctypes.data for s in seq}) == len(seq)
    with pytest.raises(ValueError):
        list(b.predict_batches([batches[0], batches[0][:, :1]]))


def test_variable_assign_and_per_block_state_reset():
    """tf.Variable.assign on a kernel changes the next forward (the packed operand copies are rebuilt);
    DownBlock2D.reset_states_per_batch (Networks.py:77-84) touches its own block only."""
    ora, model = make_pair(NET_TWO, 'NCHW', False, 9, precision='bf16x3')
    x = np.random.default_rng(3).standard_normal((2, 2, 1, 16, 16)).astype(np.float32)
    before = model(x, False)[0].numpy().copy()
    before_states = model.get_states()
    model.DownLayers[1].reset_states_per_batch(np.array([0.0, 1.0], np.float32))
    after = model.get_states()
    for lay in range(len(after[0])):
        for which in (0, 1):
            assert np.array_equal(after[0][lay][which], before_states[0][lay][which])       # level 0 untouched
    assert np.all(after[1][0][0][0] == 0) and np.all(after[1][0][1][0] == 0)
    assert np.array_equal(after[1][0][1][1], before_states[1][0][1][1])
    model.reset_states_per_batch(np.zeros(2, np.float32))
    v = [t for t in model.trainable_variables if t.name == 'UpLayers/1/Conv/1/kernel'][0]
    v.assign(v.numpy() * 2.0)
    ora.params['UpLayers/1/Conv/1/kernel'] *= 2.0
    ora.reset_states_per_batch(np.zeros(2, np.float32)) if ora.states[0][0] is not None else None
    ref = ora(torch.from_numpy(x), False)[0].numpy()
    got = model(x, False)[0].numpy()
    assert rel_err(got, ref) < 1e-3 and rel_err(got, before) > 1e-2


def test_multi_channel_image_reference_unit_test_case():
    """The reference's own unit_test input (Networks.py:266-270): 3 image channels, channels-last, 35x35, pad_image,
    successive stateful calls -- against the oracle (the soft-max axis quirk of channels-last included)."""
    from lstm_unet_b200.Networks import ULSTMnet2D
    params = O.init_params(NET_ODD, seed=21, randomize_bn=True, in_channels=3)
    ora = O.OracleNet(NET_ODD, 'NHWC', True, params=params, in_channels=3)
    model = ULSTMnet2D(NET_ODD, 'NHWC', True, precision='bf16x3')
    model.set_weights_dict({k: v.numpy().copy() for k, v in params.items()})
    rng = np.random.default_rng(2)
    for call in range(3):
        x = rng.standard_normal((2, 2, 35, 35, 3)).astype(np.float32)
        ref_l, ref_s = ora(torch.from_numpy(x), False)
        logits, softmax = model(x, False)
        assert tuple(logits.shape) == (2, 2, 35, 35, 3)
        assert rel_err(logits.numpy(), ref_l.numpy()) < 1e-3 and rel_err(softmax.numpy(), ref_s.numpy()) < 1e-3
    with pytest.raises(ValueError):
        model(np.zeros((2, 2, 35, 35, 1), np.float32), False)      # the channel count is frozen by the first call
