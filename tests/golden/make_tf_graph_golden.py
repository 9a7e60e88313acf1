"""Generates tests/golden/tf_graph_opencv.npz: outputs of the reference network's TensorFlow-op graph EXECUTED BY OPENCV.

TensorFlow itself is not installable, so neither the oracle nor the CUDA path can be compared with the
reference's own outputs.  This is the nearest thing available offline: `tests/tf_graphdef.py` writes ULSTMnet2D's forward
as the GraphDef of TensorFlow ops Keras-2 lowers it to, and OpenCV's TensorFlow importer (an implementation of those ops
that shares no code with torch or with this repository) runs it.  The vectors it produces are committed so that the GPU
tests compare the tcgen05 path with them directly.  Cases:
  * 'pad'  -- the inputs and weights of forward_pad.npz (3 levels, 21x26, pad_image, two stateful calls of T = 2 == one
              unrolled sequence of 4 frames), so the same case is covered by the fp64 oracle AND by OpenCV;
  * 'odd'  -- 4 levels, the reference unit_test's 35x35 pad_image shape (Networks.py:256-277), B = 2, T = 3.
The graph is produced by EXECUTING THE REFERENCE'S OWN, UNMODIFIED Networks.py (/root/reference) on a stand-in for the
TensorFlow / Keras names it uses whose layers emit TensorFlow ops (tests/keras_graph_standin.py): padding / crop
arithmetic, block wiring, skip order, reshapes and the stateful second call are the reference's code, the arithmetic is
OpenCV's.  `tests/tf_graphdef.build_ulstm_graph` writes the same graph without the reference (so the vectors can be
re-derived on a box that does not have it); this script asserts the two give identical outputs.
python tests/golden/make_tf_graph_golden.py      (build container only: needs cv2 and /root/reference; no TensorFlow)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import lstm_unet_oracle as O            # noqa: E402   (only for the seeded weight initialiser)
from tests import tf_graphdef as G                  # noqa: E402
from tests.golden.make_golden import NET            # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('LSTM_UNET_REFERENCE', '/root/reference')
NET_ODD = {
    'down_conv_kernels': [[(3, 8), (3, 8)], [(3, 12), (3, 12)], [(3, 12), (3, 12)], [(3, 16), (3, 16)]],
    'lstm_kernels': [[(5, 8)], [(5, 12)], [(5, 12)], [(5, 16)]],
    'up_conv_kernels': [[(3, 12), (3, 12)], [(3, 8), (3, 8)], [(3, 8), (3, 8)], [(3, 4), (3, 4), (1, 3)]],
}


def reference_driven(net, params, xs, pad_image):
    """xs: list over stateful calls of (B, T, 1, H, W) arrays.  Returns per call (logits, softmax) and the final (h, c) of every
    ConvLSTM layer, from the graph the reference's ULSTMnet2D.call emits."""
    import importlib
    from tests import keras_graph_standin as KG
    remove = KG.install()
    sys.path.insert(0, REF)
    saved = sys.modules.pop('Networks', None)
    try:
        RN = importlib.import_module('Networks')
        B, T, C, H, W = xs[0].shape
        model = RN.ULSTMnet2D(net, 'NCHW', pad_image)
        KG.load_weights(model, params)
        g = KG.Graph(B, H, W, C, T, len(xs), True)
        outs = []
        for c in range(len(xs)):
            logits, softmax = model(g.input(c), False)
            outs += [logits, softmax]
        lstm = [l for blk in model.DownLayers for l in blk.ConvLSTM]
        st = [KG.Sym(g, [n], B, 0, 0, 0, merged=False) for l in lstm for n in l.state]
        vals = KG.evaluate(g, xs, outs + st)
    finally:
        sys.modules.pop('Networks', None)
        if saved is not None:
            sys.modules['Networks'] = saved
        sys.path.remove(REF)
        remove()
    n = 2 * len(xs)
    states = [v[:, 0] for v in vals[n:]]
    return [(vals[2 * c], vals[2 * c + 1]) for c in range(len(xs))], [(states[2 * i], states[2 * i + 1]) for i in range(len(lstm))]


def main():
    out = {}
    z = np.load(os.path.join(HERE, 'forward_pad.npz'))
    params = {k[2:]: z[k] for k in z.files if k.startswith('p:')}
    x = z['x']                                                       # (call, B, T, 1, H, W)
    calls, states = reference_driven(NET, params, [x[0], x[1]], True)
    logits = np.concatenate([calls[0][0], calls[1][0]], axis=1)
    soft = np.concatenate([calls[0][1], calls[1][1]], axis=1)
    seq = np.concatenate([x[0], x[1]], axis=1)                       # two stateful calls == one sequence of 4 frames
    l2, s2, st2 = G.ulstm_forward_opencv(NET, params, seq.transpose(0, 1, 3, 4, 2), True)
    assert np.array_equal(l2, logits) and np.array_equal(s2, soft) and np.array_equal(st2[2][1], states[2][1])
    out.update({'pad:logits': logits, 'pad:softmax': soft, 'pad:h_lvl0': states[0][0], 'pad:c_lvl2': states[2][1]})

    p = {k: v.numpy() for k, v in O.init_params(NET_ODD, seed=11, randomize_bn=True).items()}
    xo = np.random.default_rng(5).standard_normal((2, 3, 1, 35, 35)).astype(np.float32)
    calls, states = reference_driven(NET_ODD, p, [xo], True)
    logits, soft = calls[0]
    l2, s2, st2 = G.ulstm_forward_opencv(NET_ODD, p, xo.transpose(0, 1, 3, 4, 2), True)
    assert np.array_equal(l2, logits) and np.array_equal(s2, soft)
    out.update({'odd:x': xo, 'odd:logits': logits, 'odd:softmax': soft})
    for i, (h, c) in enumerate(states):
        assert np.array_equal(h, st2[i][0]) and np.array_equal(c, st2[i][1])
        out['odd:h%d' % i], out['odd:c%d' % i] = h, c
    out.update({'odd:p:' + k: v for k, v in p.items()})
    np.savez_compressed(os.path.join(HERE, 'tf_graph_opencv.npz'), **out)
    print('wrote tf_graph_opencv.npz:', {k: v.shape for k, v in out.items() if ':p:' not in k})


if __name__ == '__main__':
    main()
