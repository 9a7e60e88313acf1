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


def test_reference_inference_script_equals_oracle_pipeline(ref_networks, tmp_path):
    """The reference's whole ``Inference2D.inference()`` (Inference2D.py:25-136, imported unmodified) on the stand-in:
    model_params.pickle + model.ckpt loading, the CTCInferenceReader sequence with its reversed warm-up prefix, the
    per-frame B=1 / T=1 stateful call, the post-processing and the mask TIFFs it writes -- against the oracle pipeline
    (OracleNet streaming + postprocess_frame) on the same TIFF frames.  Pins the orchestration of SURVEY row a12."""
    import pickle
    import types
    import cv2
    from oracle import postprocess_oracle as P
    from lstm_unet_b200 import tf_checkpoint
    RN, standin = ref_networks
    net = NET_B
    params = O.init_params(net, seed=21, randomize_bn=True)
    # a model directory as train2D.py leaves it (train2D.py:232-240)
    model_dir, seq_dir, out_dir = tmp_path / 'model', tmp_path / 'seq', tmp_path / 'out'
    os.makedirs(model_dir); os.makedirs(seq_dir)
    tf_checkpoint.save_model_weights(str(model_dir / 'model.ckpt'), {k: v.numpy() for k, v in params.items()})
    with open(model_dir / 'model_params.pickle', 'wb') as f:
        pickle.dump({'name': 'ULSTMnet2D', 'params': (net,)}, f)
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:24, 0:28]
    raws = []
    for t in range(4):
        a = (rng.integers(0, 80, size=(24, 28)) + 900 * (((yy - 8 - t) ** 2 + (xx - 10) ** 2) < 16)).astype(np.uint16)
        cv2.imwrite(str(seq_dir / ('t%03d.tif' % t)), a)
        raws.append(a)
    # the modules Inference2D.py imports next to Networks
    saved = {n: sys.modules.pop(n, None) for n in ('Inference2D', 'Params', 'DataHandeling', 'utils', 'distutils', 'distutils.util')}
    du = types.ModuleType('distutils.util'); du.strtobool = lambda v: int(str(v).lower() in ('1', 'true', 'y', 'yes'))
    sys.modules['distutils'] = types.ModuleType('distutils'); sys.modules['distutils.util'] = du
    # keras_standin's Model.load_weights needs the architecture: the reference passes it to the constructor
    orig_init = RN.ULSTMnet2D.__init__

    def init_and_remember(self, net_params=RN.DEFAULT_NET_DOWN_PARAMS, *a, **k):
        orig_init(self, net_params, *a, **k)
        self._standin_net_params = net_params
    RN.ULSTMnet2D.__init__ = init_and_remember
    try:
        ref_inf = importlib.import_module('Inference2D')
        assert ref_inf.__file__.startswith('/root/reference')
        pre = 2
        ref_inf.params = types.SimpleNamespace(
            model_path=str(model_dir), gpu_id=-1, data_format='NCHW', dry_run=False, save_intermediate=False,
            save_intermediate_path=None, output_path=str(out_dir), data_reader=ref_inf.DataHandeling.CTCInferenceReader,
            sequence_path=str(seq_dir), filename_format='t*.tif', pre_sequence_frames=pre, edge_dist=2, FOV=0,
            min_cell_size=1, max_cell_size=10000)
        os.makedirs(out_dir)
        ref_inf.inference()
    finally:
        RN.ULSTMnet2D.__init__ = orig_init
        for n, m in saved.items():
            sys.modules.pop(n, None)
            if m is not None:
                sys.modules[n] = m
    # the oracle pipeline on the same frames
    ora = O.OracleNet(net, 'NCHW', True, params={k: v.clone() for k, v in params.items()})
    frames = [(a.astype(np.float32) - a.astype(np.float32).mean()) / a.astype(np.float32).std() for a in raws]
    sequence = frames[:pre][::-1] + frames
    for T, img in enumerate(sequence):
        _, sm = ora(torch.from_numpy(img).reshape(1, 1, 1, 24, 28), False)
        t = T - pre
        if t < 0:
            continue
        want = P.postprocess_frame(sm[0, 0].numpy(), edge_dist=2, min_cell_size=1, max_cell_size=10000)
        got = cv2.imread(str(out_dir / ('mask%03d.tif' % t)), -1)
        assert got is not None and got.dtype == np.uint16 and np.array_equal(got, want), t
    assert sorted(os.listdir(out_dir)) == ['mask%03d.tif' % t for t in range(4)]
