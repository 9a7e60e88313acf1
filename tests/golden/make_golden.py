"""Generates tests/golden/*.npz from the CPU oracle (fp64).  The reference itself cannot produce fixtures here
(TensorFlow is not installable; SURVEY 8c "parity unpinned"), so these vectors freeze the oracle: they guard it against
drift and give the GPU tests a committed, size-stable target.   python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lstm_unet_oracle as O   # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NET = {
    'down_conv_kernels': [[(3, 8), (3, 8)], [(3, 12), (3, 12)], [(3, 16), (3, 16)]],
    'lstm_kernels': [[(5, 8)], [(5, 12)], [(3, 16)]],
    'up_conv_kernels': [[(3, 12), (3, 12)], [(3, 8), (3, 8)], [(3, 4), (3, 4), (1, 3)]],
}
CW = [0.15, 0.25, 0.6]


def forward_case():
    """pad_image inference (Inference2D call), two stateful calls, odd size (pad 4 / 4+3)."""
    p32 = O.init_params(NET, seed=101, randomize_bn=True)
    p64 = {k: v.double() for k, v in p32.items()}
    net = O.OracleNet(NET, 'NCHW', True, params=p64, dtype=torch.float64)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((2, 2, 2, 1, 21, 26)).astype(np.float32)      # (call, B, T, 1, H, W)
    out = {}
    for c in range(2):
        logits, softmax = net(torch.from_numpy(x[c]).double(), False)
        out['logits%d' % c] = logits.numpy().astype(np.float32)
        out['softmax%d' % c] = softmax.numpy().astype(np.float32)
    st = net.get_states()
    out['h_lvl0'] = st[0][0][0].astype(np.float32)
    out['c_lvl2'] = st[2][0][1].astype(np.float32)
    np.savez_compressed(os.path.join(HERE, 'forward_pad.npz'), x=x, **{'p:' + k: v.numpy() for k, v in p32.items()}, **out)


def train_case():
    """one train2D.train_step (training-mode BN, weighted CE, gradients, Keras Adam) in fp64."""
    p32 = O.init_params(NET, seed=202, randomize_bn=True)
    p64 = {k: v.double() for k, v in p32.items()}
    net = O.OracleNet(NET, 'NCHW', False, params=p64, dtype=torch.float64)
    rng = np.random.default_rng(9)
    x = rng.standard_normal((1, 2, 1, 16, 24)).astype(np.float32)
    lab = rng.integers(-1, 3, size=(1, 2, 1, 16, 24)).astype(np.float32)
    names = net.trainable_names()
    m = {n: torch.zeros_like(net.params[n]) for n in names}
    v = {n: torch.zeros_like(net.params[n]) for n in names}
    loss, logits, _, grads = O.train_step(net, torch.from_numpy(x).double(), torch.from_numpy(lab).double(), CW, m, v, 1, 1e-3)
    np.savez_compressed(os.path.join(HERE, 'train_step.npz'), x=x, lab=lab, loss=np.float64(loss),
                        logits=logits.numpy().astype(np.float32),
                        **{'p:' + k: t.numpy() for k, t in p32.items()},
                        **{'g:' + k: t.numpy().astype(np.float32) for k, t in grads.items()},
                        **{'q:' + k: net.params[k].numpy().astype(np.float32) for k in names})


if __name__ == '__main__':
    forward_case()
    train_case()
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)))
