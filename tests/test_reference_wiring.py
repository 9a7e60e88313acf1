"""The reference's own Networks.py (imported unmodified from /root/reference) executed on a torch-backed stand-in for
TensorFlow / Keras (tests/keras_standin.py) against oracle.OracleNet: pins the oracle's WIRING -- reflect-padding and crop
arithmetic for both ``pad_image`` modes and odd sizes, block order, skip order, reshapes, ``return_logits``, the soft-max
axis (incl. the channels-last quirk), stateful carry, ``reset_states_per_batch`` / ``get_states`` / ``set_states`` -- to
the reference's code.  The Keras operator semantics inside the stand-in are the oracle's own (unpinned without TF)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from oracle import lstm_unet_oracle as O

pytestmark = pytest.mark.skipif(not os.path.exists('/root/reference/Networks.py'),
                                reason='the reference is only present in the build container')

NET_A = {'down_conv_kernels': [[(3, 6), (3, 5)], [(3, 8)], [(3, 7), (3, 7)]], 'lstm_kernels': [[(3, 4)], [(5, 6), (3, 5)], [(3, 6)]],
         'up_conv_kernels': [[(3, 8)], [(3, 6), (3, 5)], [(3, 5), (1, 3)]]}
NET_B = {'down_conv_kernels': [[(3, 4)], [(3, 6)]], 'lstm_kernels': [[(5, 3)], [(3, 5)]], 'up_conv_kernels': [[(3, 4)], [(3, 4), (1, 3)]]}


@pytest.fixture()
def ref_networks():
    from tests import keras_standin
    remove = keras_standin.install()
    sys.path.insert(0, '/root/reference')
    saved = sys.modules.pop('Networks', None)
    try:
        mod = importlib.import_module('Networks')
        yield mod, keras_standin
    finally:
        sys.modules.pop('Networks', None)
        if saved is not None:
            sys.modules['Networks'] = saved
        sys.path.remove('/root/reference')
        remove()


def _pair(ref_networks, net, fmt, pad):
    RN, standin = ref_networks
    params = O.init_params(net, seed=11, randomize_bn=True)
    ref = RN.ULSTMnet2D(net, fmt, pad)
    standin.load_weights(ref, {k: v.clone() for k, v in params.items()})
    ora = O.OracleNet(net, fmt, pad, params={k: v.clone() for k, v in params.items()})
    return ref, ora


@pytest.mark.parametrize('net,fmt,pad,shape', [
    (NET_A, 'NCHW', True, (2, 3, 1, 21, 26)), (NET_A, 'NCHW', False, (2, 2, 1, 19, 24)), (NET_A, 'NCHW', False, (1, 2, 1, 16, 24)),
    (NET_B, 'NHWC', True, (2, 2, 14, 9, 1)), (NET_B, 'NWHC', False, (3, 1, 10, 12, 1))])
def test_reference_call_equals_oracle(ref_networks, net, fmt, pad, shape):
    ref, ora = _pair(ref_networks, net, fmt, pad)
    rng = np.random.default_rng(0)
    for call in range(3):                                  # stateful: later calls start from the carried h / c
        x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
        training = call == 1
        rl, rs = ref(x, training)
        ol, os_ = ora(x, training)
        assert tuple(rl.shape) == tuple(ol.shape) == shape[:2] + ((3,) + shape[3:] if fmt[1] == 'C' else shape[2:4] + (3,))
        np.testing.assert_allclose(rl.detach().numpy(), ol.detach().numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(rs.detach().numpy(), os_.detach().numpy(), rtol=1e-5, atol=1e-6)
        if call == 0:
            mask = np.array([1.0] + [0.0] * (shape[0] - 1), dtype=np.float32)
            ref.reset_states_per_batch(torch.from_numpy(mask))
            ora.reset_states_per_batch(mask)
    rstates, ostates = ref.get_states(), ora.get_states()
    for rb, ob in zip(rstates, ostates):
        for rl_, ol_ in zip(rb, ob):
            for a, b in zip(rl_, ol_):
                np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6)


def test_reference_state_swap_equals_oracle(ref_networks):
    """train2D's validation swap (train2D.py:192-220): get_states / set_states / set_states(None-filled)"""
    ref, ora = _pair(ref_networks, NET_B, 'NCHW', False)
    rng = np.random.default_rng(1)
    x1 = torch.from_numpy(rng.standard_normal((2, 2, 1, 8, 12)).astype(np.float32))
    x2 = torch.from_numpy(rng.standard_normal((2, 2, 1, 8, 12)).astype(np.float32))
    fresh_r, fresh_o = ref.get_states(), ora.get_states()             # None before the first call
    assert fresh_r[0][0][0] is None and fresh_o[0][0][0] is None
    ref(x1, False); ora(x1, False)
    saved_r, saved_o = ref.get_states(), ora.get_states()
    ref.set_states(fresh_r); ora.set_states(fresh_o)                  # back to "no state"
    a, _ = ref(x2, False); b, _ = ora(x2, False)
    np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-5, atol=1e-6)
    ref.set_states(saved_r); ora.set_states(saved_o)                  # restore the training states
    a, _ = ref(x2, False); b, _ = ora(x2, False)
    np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('channel_axis,shape', [(2, (2, 3, 3, 9, 11)), (4, (2, 2, 7, 8, 3))])
def test_reference_weighted_ce_loss_equals_oracle(ref_networks, channel_axis, shape):
    """losses.WeightedCELoss.__call__ (losses.py:13-27), imported unmodified, on the stand-in's tensor ops against the
    oracle's weighted_ce_loss: ignore label -1, class weights, normalisation by the valid pixels + 1e-5, both layouts"""
    saved = sys.modules.pop('losses', None)
    try:
        ref_losses = importlib.import_module('losses')
        assert ref_losses.__file__.startswith('/root/reference')
        rng = np.random.default_rng(2)
        logits = torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
        lab_shape = list(shape)
        lab_shape[channel_axis] = 1
        labels = torch.from_numpy(rng.integers(-1, 3, size=lab_shape).astype(np.float32))
        cw = [0.15, 0.25, 0.6]
        want = ref_losses.WeightedCELoss(channel_axis, cw)(labels, logits)
        got = O.weighted_ce_loss(labels, logits, cw, channels_first=channel_axis == 2)
        assert abs(float(want) - float(got)) < 1e-6 * max(1.0, abs(float(want)))
        none_valid = torch.full(lab_shape, -1.0)
        assert float(ref_losses.WeightedCELoss(channel_axis, cw)(none_valid, logits)) == 0.0
        assert float(O.weighted_ce_loss(none_valid, logits, cw, channels_first=channel_axis == 2)) == 0.0
    finally:
        sys.modules.pop('losses', None)
        if saved is not None:
            sys.modules['losses'] = saved
