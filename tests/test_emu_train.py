"""Host-logic parity of the training step (loss, backward, Adam) through the C-ABI, TEST-ONLY host build, vs autograd
through the oracle.  Same plan / tables drive the GPU path (tests -m gpu)."""
import numpy as np
import torch

from oracle import lstm_unet_oracle as O
from tests.emu_backend import emu_session, emu_forward

NET_B = {
    'down_conv_kernels': [[(3, 6)], [(3, 10), (3, 10)]],
    'lstm_kernels': [[(3, 5), (3, 7)], [(5, 9)]],
    'up_conv_kernels': [[(3, 6)], [(3, 5), (1, 3)]],
}
NET_C = {   # 3 levels: two stride-2 convs, two up-samplings, 5x5 lstm on the image
    'down_conv_kernels': [[(3, 4), (3, 4)], [(3, 6)], [(3, 8), (3, 8)]],
    'lstm_kernels': [[(5, 4)], [(3, 6)], [(3, 8)]],
    'up_conv_kernels': [[(3, 6), (3, 6)], [(3, 4)], [(3, 4), (3, 4), (1, 3)]],
}
CW = [0.15, 0.25, 0.6]


def grad_rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


class TieWatch:
    """LeakyReLU and hard_sigmoid have kinks: a pre-activation within rounding distance of one makes the (sub)gradient
    of a single element differ by a factor 0.3 / 1.0 between two correct implementations.  The parity check is only
    meaningful on inputs without such near-ties, so the oracle's pre-activations are watched and a step with a
    near-tie is not compared."""

    def __init__(self):
        self.min_lrelu = self.min_gate = 1e9
        self._l, self._g = O.leaky_relu, O.hard_sigmoid

    def __enter__(self):
        def lrelu(x):
            self.min_lrelu = min(self.min_lrelu, float(x.detach().abs().min()))
            return self._l(x)

        def gate(x):
            self.min_gate = min(self.min_gate, float((x.detach().abs() - 2.5).abs().min()))
            return self._g(x)
        O.leaky_relu, O.hard_sigmoid = lrelu, gate
        return self

    def __exit__(self, *a):
        O.leaky_relu, O.hard_sigmoid = self._l, self._g

    def clean(self, tol=2e-5):
        return self.min_lrelu > tol and self.min_gate > tol


def run_case(net, B, T, H, W, seed, steps=2, smooth=False, C=1):
    params = O.init_params(net, seed=seed, randomize_bn=True, in_channels=C)
    p_np = {k: v.numpy().copy() for k, v in params.items()}
    ora = O.OracleNet(net, 'NCHW', False, params=params, gate='sigmoid' if smooth else 'hard_sigmoid')
    if not smooth:
        ora.gate = lambda x: O.hard_sigmoid(x)       # late-bound so TieWatch sees the calls
    sess = emu_session(net, data_format='NCHW', pad_image=False, batch=B, max_t=T, height=H, width=W,
                       precision='bf16x3', train=True, gate='sigmoid' if smooth else 'hard_sigmoid',
                       lrelu_alpha=1.0 if smooth else 0.3, in_channels=C)
    sess.set_params(p_np)
    names = ora.trainable_names()
    m = {n: torch.zeros_like(ora.params[n]) for n in names}
    v = {n: torch.zeros_like(ora.params[n]) for n in names}
    nt = sess.n_trainable
    grads = np.zeros(nt, dtype=np.float32)
    am = np.zeros(nt, dtype=np.float32)
    av = np.zeros(nt, dtype=np.float32)
    loss = np.zeros(1, dtype=np.float32)
    rng = np.random.default_rng(seed)
    lr = 1e-3
    n_compared = 0
    for step in range(1, steps + 1):          # second step: non-zero initial h/c (truncated BPTT) and updated weights
        x = rng.standard_normal((B, T, C, H, W)).astype(np.float32)
        lab = rng.integers(-1, 3, size=(B, T, 1, H, W)).astype(np.float32)
        with TieWatch() as tw:
            ref_loss, ref_logits, _, ref_grads = O.train_step(ora, torch.from_numpy(x), torch.from_numpy(lab), CW, m, v, step, lr)
        logits, _ = emu_forward(sess, x, True)
        assert grad_rel(logits, ref_logits.numpy()) < 1e-3
        sess.loss_backward(lab.ctypes.data, CW, loss.ctypes.data, grads.ctypes.data)
        assert abs(float(loss[0]) - float(ref_loss)) < 1e-4 * max(1.0, abs(float(ref_loss)))
        worst = ('', 0.0)
        compared = tw.clean() or smooth
        for e in sess.layout:
            if not compared:
                break
            if not e['trainable']:
                continue
            g = grads[e['offset']:e['offset'] + e['count']].reshape(e['shape'])
            r = ref_grads[e['name']].numpy()
            scale = max(np.abs(r).max(), 1e-6)
            err = float(np.abs(g - r).max() / scale)
            # conv biases feeding a BatchNorm have an analytically zero gradient: compare absolutely
            if e['name'].endswith('bias') and 'ConvLSTM' not in e['name'] and np.abs(r).max() < 1e-6:
                assert np.abs(g).max() < 1e-4, (e['name'], np.abs(g).max())
                continue
            if err > worst[1]:
                worst = (e['name'], err)
            assert err < 5e-3, (step, e['name'], err)
        sess.adam_step(grads.ctypes.data, am.ctypes.data, av.ctypes.data, lr, step)
        got = sess.get_params()
        for n in (names if compared else []):
            if n.endswith('bias') and 'ConvLSTM' not in n and 'UpLayers/%d/Conv/%d' % (len(net['up_conv_kernels']) - 1, len(net['up_conv_kernels'][-1]) - 1) not in n:
                continue                      # zero-gradient biases: Adam amplifies rounding noise to +-lr
            d = np.abs(got[n] - ora.params[n].numpy()).max()
            assert d < 0.2 * lr + 1e-6, (step, n, d)
        n_compared += 1 if compared else 0
    sess.close()
    return worst, n_compared


def test_train_step_two_levels():
    worst, n = run_case(NET_B, 2, 2, 8, 8, 23)
    assert n == 2, 'choose a seed without near-ties'


def test_train_step_three_levels():
    worst, n = run_case(NET_C, 1, 2, 16, 16, 31)
    assert n == 2, 'choose a seed without near-ties'


def test_train_step_smooth_variant_needs_no_tie_watch(monkeypatch):
    """sigmoid gates + LeakyReLU slope 1 (lu_config.gate / lrelu_alpha): the gradient is continuous in the forward, every
    step is compared -- the configuration tests/test_gpu_ctc_parity.py uses for the CTC-size backward.  Seed 21 has a
    near-tie in the reference configuration (next test)."""
    monkeypatch.setattr(O, 'LRELU_ALPHA', 1.0)
    worst, n = run_case(NET_B, 2, 2, 8, 8, 21, smooth=True)
    assert n == 2 and worst[1] < 1e-3


def test_train_step_multi_channel_image(monkeypatch):
    """in_channels = 3: weight gradients of the layers that read the image (no data gradient flows into it)."""
    monkeypatch.setattr(O, 'LRELU_ALPHA', 1.0)
    worst, n = run_case(NET_C, 1, 2, 16, 16, 5, smooth=True, C=3)
    assert n == 2 and worst[1] < 1e-3


def test_near_tie_step_is_detected():
    # seed 21 puts one LeakyReLU pre-activation within 1e-6 of zero at step 2 (a single-element 0.3-vs-1.0 subgradient
    # difference); the watch must flag it
    worst, n = run_case(NET_B, 2, 2, 8, 8, 21)
    assert n == 1


def test_loss_only_matches_oracle():
    net, B, T, H, W = NET_B, 2, 1, 8, 8
    params = O.init_params(net, seed=5, randomize_bn=True)
    ora = O.OracleNet(net, 'NCHW', False, params=params)
    sess = emu_session(net, data_format='NCHW', pad_image=False, batch=B, max_t=T, height=H, width=W, precision='bf16x3')
    sess.set_params({k: v.numpy().copy() for k, v in params.items()})
    rng = np.random.default_rng(1)
    x = rng.standard_normal((B, T, 1, H, W)).astype(np.float32)
    lab = rng.integers(-1, 3, size=(B, T, 1, H, W)).astype(np.float32)
    ref_logits, _ = ora(torch.from_numpy(x), False)
    ref = O.weighted_ce_loss(torch.from_numpy(lab), ref_logits, CW, True)
    emu_forward(sess, x, False)
    loss = np.zeros(1, dtype=np.float32)
    sess.loss_backward(lab.ctypes.data, CW, loss.ctypes.data, None)
    assert abs(float(loss[0]) - float(ref)) < 1e-4
    sess.close()
