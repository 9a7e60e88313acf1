"""Task lists of the tcgen05 weight-gradient kernels, replayed on the host (TEST-ONLY build, LU_WGRAD_EMU_TASKS) and compared
with the scalar mirror on the same forward pass and the same upstream gradients.

The replay (emulate_wg_tasks / emulate_wg_pair_tasks in csrc/lu_train_host.inl) does, with scalar loops, what
lu_wgrad_tc_kernel / lu_wgrad_pair_kernel do with a task: halo windows with the tensor map's zero fill, taps as row offsets
into them, 128-pixel tiles, one accumulator per tap, the flush into the packed gradient; for a pair task the transposed
product (each CTA: its 128 output channels x the input-channel chunks of both CTAs).  LU_WGRAD_EMU_TASKS = the
LU_WGRAD_PAIR setting replayed: 0 = independent CTAs only, 1 = CTA pairs with one input chunk per CTA, 2 = two chunks per
CTA where a source has four, 3 = one chunk per CTA with tap-pair accumulator entries (N = 256).  What this pins on the CPU is the task builder: every (stage, tap, column chunk, pixel range)
exactly once across the two lists, K block indices, which columns / chunks fall back to independent tasks.  The
instruction-level protocol of the kernels is covered by the -m gpu tests."""
import os

import numpy as np
import pytest

from oracle import lstm_unet_oracle as O
from tests.emu_backend import emu_session, emu_forward

# >= 65 output channels somewhere, so that two-chunk column slabs occur; stride-2 convs
# (tap subsets per parity plane), a 5x5 and a 3x3 ConvLSTM, a 1x1 conv
NET_W = {
    'down_conv_kernels': [[(3, 70)], [(3, 66), (3, 20)]],
    'lstm_kernels': [[(5, 33)], [(3, 40)]],
    'up_conv_kernels': [[(3, 68)], [(3, 6), (1, 3)]],
}
# sources with 2 and with 4 chunks, 4 and 8 output chunks: pair tasks with one and with two chunks per CTA, remainder
# columns and odd chunks left to independent tasks
NET_P = {
    'down_conv_kernels': [[(3, 200)], [(3, 196), (3, 20)]],
    'lstm_kernels': [[(3, 70)], [(3, 66)]],
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
        if mode is not None:
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


@pytest.mark.parametrize('mode', [0, 1, 2, 3])
def test_task_replay_equals_scalar_mirror_bf16(mode):
    """bf16 mode, wide net (few layers can pair: most sources have one chunk)"""
    ref, layout = grads_of(NET_W, 'bf16', None)
    got, _ = grads_of(NET_W, 'bf16', mode)
    # same bf16 operands, fp32 accumulation in a different order
    compare(ref, got, layout, 2e-5)


@pytest.mark.parametrize('mode', [1, 2, 3])
def test_pair_task_replay_equals_scalar_mirror(mode, capfd):
    """a net whose layers DO pair: 2- and 4-chunk sources, 4 and 8 output chunks"""
    os.environ['LU_WGRAD_EMU_VERBOSE'] = '1'
    try:
        ref, layout = grads_of(NET_P, 'bf16', None, H=8)
        got, _ = grads_of(NET_P, 'bf16', mode, H=8)
    finally:
        os.environ.pop('LU_WGRAD_EMU_VERBOSE', None)
    compare(ref, got, layout, 2e-5)
    err = capfd.readouterr().err
    import re
    pairs = sum(int(m) for m in re.findall(r': (\d+) pair', err))
    assert pairs > 0, 'no pair task was built for a net that should pair'


@pytest.mark.parametrize('mode', [0, 2])
def test_task_replay_equals_scalar_mirror_bf16x3(mode):
    """the parity mode (hi / lo planes: activation-lo tasks pair with the hi plane of dY only); a pair request falls back
    to independent tasks"""
    ref, layout = grads_of(NET_S, 'bf16x3', None, B=2, H=8)
    got, _ = grads_of(NET_S, 'bf16x3', mode, B=2, H=8)
    compare(ref, got, layout, 2e-5)
