"""Stand-alone DownBlock2D / UpBlock2D handles (the reference's unit_test usage, Networks.py:100-119,155-175): host build
of the library (plan, tables, packing, epilogues -- scalar mirror engine) against the block oracle."""
import numpy as np
import pytest
import torch

from lstm_unet_b200 import _lib
from lstm_unet_b200.session import LuError
from oracle import blocks_oracle as BO
from tests.emu_backend import emu_block_forward, emu_block_session

TOL = 1e-3          # the north_star tolerance; bf16x3 measures ~1e-5


def _rel(a, b):
    b = np.asarray(b)
    return float(np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-30))


def _np(params):
    return {k: v.numpy() for k, v in params.items()}


@pytest.mark.parametrize('data_format,stride', [('NCHW', 2), ('NHWC', 2), ('NCHW', 1)])
def test_down_block_alone_matches_oracle(data_format, stride):
    conv_kernels, lstm_kernels = [(3, 16), (3, 24)], [(3, 12), (5, 16)]
    B, T, H, W, C = 2, 3, 12, 16, 3                               # the reference unit_test feeds 3 channels
    ora = BO.OracleDownBlock(conv_kernels, lstm_kernels, stride, data_format, in_channels=C, seed=3)
    cfg = _lib.make_down_block_config(conv_kernels, lstm_kernels, stride, data_format, batch=B, max_t=T, height=H, width=W,
                                      in_channels=C, precision='bf16x3', engine='simt')
    sess = emu_block_session(cfg)
    assert [e['name'] for e in sess.layout if e['trainable']][:3] == [
        'DownLayers/0/ConvLSTM/0/kernel', 'DownLayers/0/ConvLSTM/0/recurrent_kernel', 'DownLayers/0/ConvLSTM/0/bias']
    assert not any(e['name'].startswith('UpLayers') for e in sess.layout)
    sess.set_params(_np(ora.params))
    rng = np.random.default_rng(0)
    shape = (B, T, C, H, W) if data_format == 'NCHW' else (B, T, H, W, C)
    for call, training in enumerate((False, False, True, False)):  # stateful carry, BatchNorm batch statistics, new moving statistics
        x = rng.standard_normal(shape).astype(np.float32)
        down, activ = ora(torch.from_numpy(x), training)
        got = emu_block_forward(sess, x, training=training)
        assert got.shape == tuple(activ.shape)
        assert tuple(down.shape[2:]) == got.shape[1:] and down.shape[0] * down.shape[1] == got.shape[0]
        assert _rel(got, activ.numpy()) < TOL, (call, _rel(got, activ.numpy()))
        np.testing.assert_allclose(got.reshape(down.shape), down.numpy(), atol=TOL * float(np.abs(activ.numpy()).max()))
    # moving statistics moved like the oracle's (training call above)
    mv = sess.get_params()['DownLayers/0/BN/1/moving_variance']
    assert _rel(mv, ora.params['DownLayers/0/BN/1/moving_variance'].numpy()) < TOL
    sess.close()


@pytest.mark.parametrize('data_format,up_factor,return_logits', [('NCHW', 2, False), ('NHWC', 2, True), ('NCHW', 1, True)])
def test_up_block_alone_matches_oracle(data_format, up_factor, return_logits):
    kernels = [(3, 16), (3, 8), (1, 3)]
    N, h, w, C, Cs = 3, 6, 10, 5, 3
    ora = BO.OracleUpBlock(kernels, up_factor, data_format, return_logits, in_channels=C, skip_channels=Cs, seed=5)
    cfg = _lib.make_up_block_config(kernels, up_factor, data_format, return_logits, frames=N, height=h, width=w,
                                    in_channels=C, skip_channels=Cs, precision='bf16x3', engine='simt')
    sess = emu_block_session(cfg)
    names = [e['name'] for e in sess.layout]
    assert ('UpLayers/0/BN/2/gamma' in names) == (not return_logits)        # Networks.py:148-149
    sess.set_params(_np(ora.params))
    rng = np.random.default_rng(1)
    H, W = h * up_factor, w * up_factor
    for training in (False, True, False):
        x = rng.standard_normal((N, C, h, w) if data_format == 'NCHW' else (N, h, w, C)).astype(np.float32)
        skip = rng.standard_normal((N, Cs, H, W) if data_format == 'NCHW' else (N, H, W, Cs)).astype(np.float32)
        ref = ora((torch.from_numpy(x), torch.from_numpy(skip)), training).numpy()
        got = emu_block_forward(sess, x, skip, training=training)
        assert got.shape == ref.shape == ((N, 3, H, W) if data_format == 'NCHW' else (N, H, W, 3))
        assert _rel(got, ref) < TOL, _rel(got, ref)
    sess.close()


def test_block_handle_errors():
    with pytest.raises(LuError, match='even input sizes'):
        emu_block_session(_lib.make_down_block_config([(3, 8)], [(3, 8)], 2, 'NCHW', batch=1, max_t=1, height=11, width=12,
                                                      in_channels=1, engine='simt'))
    with pytest.raises(LuError, match='must be 1 or 2'):
        emu_block_session(_lib.make_down_block_config([(3, 8)], [(3, 8)], 3, 'NCHW', batch=1, max_t=1, height=12, width=12,
                                                      in_channels=1, engine='simt'))
    sess = emu_block_session(_lib.make_down_block_config([(3, 8)], [(3, 8)], 2, 'NCHW', batch=1, max_t=1, height=12,
                                                         width=12, in_channels=1, engine='simt'))
    x = np.zeros((1, 1, 1, 12, 12), np.float32)
    out = np.zeros((1, 8, 6, 6), np.float32)
    with pytest.raises(LuError, match='lu_block_forward'):
        sess.forward(x.ctypes.data, 1, False, out.ctypes.data, out.ctypes.data)
    with pytest.raises(LuError, match='skip input belongs to UpBlock2D'):
        sess.block_forward(x.ctypes.data, x.ctypes.data, 1, False, out.ctypes.data)
    sess.close()
