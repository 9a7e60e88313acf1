"""-m gpu: DownBlock2D / UpBlock2D constructed and called on their own (the reference's unit_test usage,
Networks.py:100-119,155-175) run the tcgen05 kernels on a handle of their own; compared with the block oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    b = np.asarray(b)
    return float(np.abs(np.asarray(a) - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize('data_format,precision,tol', [('NHWC', 'bf16x3', 1e-3), ('NCHW', 'bf16', 5e-2)])
def test_down_block_unit_test_shapes_and_values(data_format, precision, tol):
    from lstm_unet_b200.Networks import DownBlock2D
    from oracle import blocks_oracle as BO
    conv_kernels = [(3, 16), (3, 32), (3, 64)]                   # Networks.py:102-103
    lstm_kernels = [(3, 16), (3, 32), (3, 64)]
    B, T, H, W, C = 2, 3, 50, 50, 3
    ora = BO.OracleDownBlock(conv_kernels, lstm_kernels, 2, data_format, in_channels=C, seed=11)
    blk = DownBlock2D(conv_kernels, lstm_kernels, 2, data_format, precision=precision)
    blk.set_weights_dict({k: v.numpy() for k, v in ora.params.items()})
    rng = np.random.default_rng(0)
    shape = (B, T, H, W, C) if data_format == 'NHWC' else (B, T, C, H, W)
    for call, training in enumerate((False, True, False)):
        x = rng.standard_normal(shape).astype(np.float32)
        down_ref, activ_ref = ora(torch.from_numpy(x), training)
        down, activ = blk(x, training)
        assert tuple(down.shape) == tuple(down_ref.shape) and tuple(activ.shape) == tuple(activ_ref.shape)
        assert tuple(activ.shape) == ((B * T, 25, 25, 64) if data_format == 'NHWC' else (B * T, 64, 25, 25))   # 50 -> 25
        assert _rel(activ.numpy(), activ_ref.numpy()) < tol, (call, _rel(activ.numpy(), activ_ref.numpy()))
        assert _rel(down.numpy(), down_ref.numpy()) < tol
    # state API of the block on its own (Networks.py:77-98)
    st = blk.get_states()
    assert len(st) == 3 and st[0][0].shape == ((B, 50, 50, 16) if data_format == 'NHWC' else (B, 16, 50, 50))
    h_ref = ora.states[2][0].numpy() if data_format == 'NCHW' else ora.states[2][0].permute(0, 2, 3, 1).numpy()
    assert _rel(st[2][0], h_ref) < tol
    blk.reset_states_per_batch(np.array([1.0, 0.0], np.float32))
    ora.reset_states_per_batch(np.array([1.0, 0.0], np.float32))
    st2 = blk.get_states()
    assert np.abs(st2[0][1][1]).max() == 0 and np.array_equal(st2[0][1][0], st[0][1][0])
    blk.set_states(st)
    x = rng.standard_normal(shape).astype(np.float32)
    for s_, o_ in zip(st, ora.states):                         # put the oracle back to the same states
        for w in (0, 1):
            t = torch.from_numpy(s_[w])
            o_[w] = t if data_format == 'NCHW' else t.permute(0, 3, 1, 2).contiguous()
    assert _rel(blk(x, False)[1].numpy(), ora(torch.from_numpy(x), False)[1].numpy()) < tol
    blk.close()


@pytest.mark.parametrize('data_format,up_factor,return_logits', [('NHWC', 2, False), ('NCHW', 2, True), ('NCHW', 1, False)])
def test_up_block_unit_test_shapes_and_values(data_format, up_factor, return_logits):
    from lstm_unet_b200.Networks import UpBlock2D
    from oracle import blocks_oracle as BO
    kernels = [(3, 16), (3, 32), (3, 64)]                        # Networks.py:157
    N, h, w, C = 6, 50, 50, 3
    ora = BO.OracleUpBlock(kernels, up_factor, data_format, return_logits, in_channels=C, skip_channels=C, seed=13)
    blk = UpBlock2D(kernels, up_factor, data_format, return_logits, precision='bf16x3')
    blk.set_weights_dict({k: v.numpy() for k, v in ora.params.items()})
    rng = np.random.default_rng(2)
    H, W = h * up_factor, w * up_factor
    for training in (True, False, True, False):     # first call in training mode (the reference's unit_test loop)
        x = rng.standard_normal((N, h, w, C) if data_format == 'NHWC' else (N, C, h, w)).astype(np.float32)
        skip = rng.standard_normal((N, H, W, C) if data_format == 'NHWC' else (N, C, H, W)).astype(np.float32)
        ref = ora((torch.from_numpy(x), torch.from_numpy(skip)), training).numpy()
        got = blk((x, skip), training)
        assert tuple(got.shape) == ref.shape == ((N, H, W, 64) if data_format == 'NHWC' else (N, 64, H, W))
        assert _rel(got.numpy(), ref) < 1e-3, _rel(got.numpy(), ref)
    with pytest.raises(ValueError):
        blk((x, skip[:, :-2] if data_format == 'NHWC' else skip[:, :, :-2]), False)
    blk.close()


def test_reference_unit_tests_run():
    """DownBlock2D.unit_test / UpBlock2D.unit_test (Networks.py:100-119,155-175): the shapes they print."""
    from lstm_unet_b200.Networks import DownBlock2D, UpBlock2D
    down, activ = DownBlock2D.unit_test()
    assert tuple(down.shape) == (2, 3, 25, 25, 64) and tuple(activ.shape) == (6, 25, 25, 64)
    assert tuple(UpBlock2D.unit_test().shape) == (6, 100, 100, 64)
    assert np.isfinite(activ.numpy()).all()


def test_down_block_unroll_growth_keeps_weights_and_states():
    """A longer unroll than the first call's re-opens the block's handle; weights and recurrent states carry over (the
    stateful ConvLSTM state is per block, Networks.py:48-50), and a different B / H / W is refused like Keras does."""
    from lstm_unet_b200.Networks import DownBlock2D
    from oracle import blocks_oracle as BO
    conv_kernels, lstm_kernels = [(3, 24)], [(5, 20)]
    ora = BO.OracleDownBlock(conv_kernels, lstm_kernels, 1, 'NCHW', in_channels=1, seed=17)
    blk = DownBlock2D(conv_kernels, lstm_kernels, 1, 'NCHW', precision='bf16x3')
    blk.set_weights_dict({k: v.numpy() for k, v in ora.params.items()})
    rng = np.random.default_rng(5)
    for T in (2, 1, 4, 3):
        x = rng.standard_normal((2, T, 1, 21, 19)).astype(np.float32)          # odd sizes are fine at stride 1
        ref = ora(torch.from_numpy(x), False)[1].numpy()
        got = blk(x, False)[1].numpy()
        assert _rel(got, ref) < 1e-3, (T, _rel(got, ref))
    with pytest.raises(ValueError):
        blk(np.zeros((3, 1, 1, 21, 19), np.float32), False)
    with pytest.raises(ValueError):
        DownBlock2D(conv_kernels, lstm_kernels, 2, 'NCHW')(np.zeros((1, 1, 1, 21, 19), np.float32), False)   # stride 2, odd size
    blk.close()
