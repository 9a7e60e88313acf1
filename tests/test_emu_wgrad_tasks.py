"""Task lists of the tcgen05 weight-gradient kernel, replayed on the host (TEST-ONLY build, LU_WGRAD_EMU_TASKS) and compared
with the scalar mirror on the same forward pass and the same upstream gradients.

The replay (emulate_wg_tasks in csrc/lu_train_host.inl) does, with scalar loops, what lu_wgrad_tc_kernel does with a task:
halo windows with the tensor map's zero fill, taps as row offsets into them, 128-pixel tiles, one accumulator per tap, the
flush into the packed gradient.  Mode 1 = independent CTAs (the product default), 2 = multicast pairs, 3 = one M = 256 MMA
per CTA pair (the odd task's window displaced by the tap difference, the even task's offsets read in both).  What this
pins on the CPU is the task builder: every (stage, tap, column chunk, pixel range) exactly once, K block indices, the tap
pairing and displacement of mode 3.  The instruction-level protocol of the kernel is covered by the -m gpu tests."""
import os

import numpy as np
import pytest

from oracle import lstm_unet_oracle as O
from tests.emu_backend import emu_session, emu_forward

# >= 65 output channels somewhere, so that two-chunk column slabs (the only ones mode 3 pairs) occur; stride-2 convs
# (tap subsets per parity plane), a 5x5 and a 3x3 ConvLSTM, a 1x1 conv
NET_W = {
    'down_conv_kernels': [[(3, 70)], [(3, 66), (3, 20)]],
    'lstm_kernels': [[(5, 33)], [(3, 40)]],
    'up_conv_kernels': [[(3, 68)], [(3, 6), (1, 3)]],
}
NET_S = {
    'down_conv_kernels': [[(3, 6)], [(3, 10), (3, 10)]],
    'lstm_kernels': [[(3, 5), (3, 7)], [(5, 9)]],
    'up_conv_kernels': [[(3, 6)], [(3, 5), (1, 3)]],
}
CW = [0.15, 0.25, 0.6]


def grads_of(net, precision, mode, B=1, T=2, H=16, W=8, seed=3):
    params = O.init_params(net, seed=seed, randomize_bn=True)
    sess = emu_session(net, data_format='NCHW', pad_image=False, batch=B, max_t=T, height=H, width=W,
                       precision=precision, train=True)
    sess.set_params({k: v.numpy().copy() for k, v in params.items()})
    rng = np.random.default_rng(seed)
    grads = np.zeros(sess.n_trainable, dtype=np.float32)
    loss = np.zeros(1, dtype=np.float32)
    out = []
    old = os.environ.pop('LU_WGRAD_EMU_TASKS', None)
    try:
        if mode:
            os.environ['LU_WGRAD_EMU_TASKS'] = str(mode)
        for step in range(2):        # second step: non-zero initial h (the t == 0 pass against the state buffer)
            x = rng.standard_normal((B, T, 1, H, W)).astype(np.float32)
            lab = rng.integers(-1, 3, size=(B, T, 1, H, W)).astype(np.float32)
            emu_forward(sess, x, True)
            sess.loss_backward(lab.ctypes.data, CW, loss.ctypes.data, grads.ctypes.data)
            out.append(grads.copy())
    finally:
        os.environ.pop('LU_WGRAD_EMU_TASKS', None)
        if old is not None:
            os.environ['LU_WGRAD_EMU_TASKS'] = old
    layout = [dict(e) for e in sess.layout]
    sess.close()
    return out, layout


def compare(ref, got, layout, tol):
    for step, (r, g) in enumerate(zip(ref, got)):
        for e in layout:
            if not e['trainable'] or not e['name'].endswith('kernel'):
                continue
            a = r[e['offset']:e['offset'] + e['count']]
            b = g[e['offset']:e['offset'] + e['count']]
            scale = max(float(np.abs(a).max()), 1e-12)
            err = float(np.abs(a - b).max()) / scale
            assert err < tol, (step, e['name'], err)


@pytest.mark.parametrize('mode', [1, 2, 3])
def test_task_replay_equals_scalar_mirror_bf16(mode):
    """bf16 mode, wide net: all three cluster modes; mode 3 pairs taps in the layers with an even number of column chunks
    and falls back to independent CTAs in the others."""
    ref, layout = grads_of(NET_W, 'bf16', 0)
    got, _ = grads_of(NET_W, 'bf16', mode)
    # same bf16 operands, fp32 accumulation in a different order
    compare(ref, got, layout, 2e-5)


@pytest.mark.parametrize('mode', [1, 2])
def test_task_replay_equals_scalar_mirror_bf16x3(mode):
    """the parity mode (hi / lo planes: activation-lo tasks pair with the hi plane of dY only)"""
    ref, layout = grads_of(NET_S, 'bf16x3', 0, B=2, H=8)
    got, _ = grads_of(NET_S, 'bf16x3', mode, B=2, H=8)
    compare(ref, got, layout, 2e-5)


def test_mode3_request_in_parity_mode_falls_back():
    ref, layout = grads_of(NET_S, 'bf16x3', 0, B=2, H=8)
    got, _ = grads_of(NET_S, 'bf16x3', 3, B=2, H=8)
    compare(ref, got, layout, 2e-5)
