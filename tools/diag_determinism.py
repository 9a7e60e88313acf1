"""Run-to-run determinism of one training step on fresh models (same weights, same inputs): forward logits bitwise?
gradients per tensor?  Usage: python tools/diag_determinism.py [runs]   (env switches LU_PAIR / LU_WGRAD_PAIR apply)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge      # noqa: E402

ge.build()
from lstm_unet_b200.Networks import ULSTMnet2D      # noqa: E402
from oracle import lstm_unet_oracle as O            # noqa: E402

NET = {'down_conv_kernels': [[(3, 64), (3, 64)], [(3, 128), (3, 128)]], 'lstm_kernels': [[(5, 64)], [(5, 128)]],
       'up_conv_kernels': [[(3, 64), (3, 64)], [(3, 32), (3, 32), (1, 3)]]}
CW = [0.15, 0.25, 0.6]
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
precision = sys.argv[2] if len(sys.argv) > 2 else 'bf16x3'
params = O.init_params(NET, seed=3, randomize_bn=True)
weights = {k: v.numpy().copy() for k, v in params.items()}
rng = np.random.default_rng(0)
x = rng.standard_normal((2, 2, 1, 32, 48)).astype(np.float32)
lab = rng.integers(-1, 3, size=(2, 2, 1, 32, 48)).astype(np.float32)
names = ['DownLayers/0/ConvLSTM/0', 'DownLayers/0/Conv/0', 'DownLayers/0/Conv/1', 'DownLayers/1/ConvLSTM/0', 'DownLayers/1/Conv/0',
         'DownLayers/1/Conv/1', 'UpLayers/0/Conv/0', 'UpLayers/0/Conv/1', 'UpLayers/1/Conv/0', 'UpLayers/1/Conv/1']
res = []
for r in range(runs):
    m = ULSTMnet2D(NET, 'NCHW', False, precision=precision, train=True)
    m.set_weights_dict(weights)
    lg, _ = m(x, True)
    acts = {n: m._sess.debug_buffer(n, 0) for n in names}
    loss, g = m.backward(lab, CW)
    gacts = {n: m._sess.debug_buffer(n, 1) for n in names}
    res.append((lg.numpy().copy(), acts, g.cpu().numpy().copy(), gacts, [dict(e) for e in m._sess.layout]))
    m.close()
a = res[0]
for r in range(1, runs):
    b = res[r]
    print('run %d vs 0: logits bitwise %s' % (r, np.array_equal(a[0], b[0])))
    for n in names:
        if not np.array_equal(a[1][n], b[1][n]):
            print('   forward activation differs first at', n, float(np.abs(a[1][n] - b[1][n]).max()))
            break
    for n in reversed(names):
        if not np.array_equal(a[3][n], b[3][n]):
            d = np.abs(a[3][n] - b[3][n])
            print('   gradient buffer differs (walking back from the loss) first at', n, 'max', float(d.max()), 'of', float(np.abs(a[3][n]).max()),
                  'elements', int((d > 0).sum()))
            break
    worst = ('', 0.0)
    for e in a[4]:
        if not e['trainable']:
            continue
        u, v = a[2][e['offset']:e['offset'] + e['count']], b[2][e['offset']:e['offset'] + e['count']]
        sc = np.abs(u).max()
        if sc < 1e-6:
            continue
        err = float(np.abs(u - v).max() / sc)
        if err > worst[1]:
            worst = (e['name'], err)
    print('   worst gradient tensor', worst)
