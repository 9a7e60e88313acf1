"""TEST-ONLY stand-in for the TensorFlow / Keras names of the reference's Networks.py that EMITS A TENSORFLOW GRAPH.

Sister of tests/keras_standin.py (which backs the same names with torch arithmetic).  Here the tensors are symbolic: every
Keras layer / tf function the reference calls appends the TensorFlow ops Keras-2 lowers it to (tests/tf_graphdef.py) to a
GraphDef, so that running the reference's OWN, unmodified ``ULSTMnet2D.call`` (/root/reference/Networks.py: pad / crop
arithmetic, block wiring, skip order, reshapes, return_logits, stateful ConvLSTM layers called twice) produces the graph,
and OpenCV's TensorFlow importer produces the numbers.  Neither the wiring nor the arithmetic of the vectors made this way
comes from this repository; what does is the layer -> op lowering below (Keras-2 defaults).

Keras-2 on a CPU runs channels_first models through NHWC kernels (its TF backend transposes around every conv / pool when
there is no NCHW support), so the ops are emitted in NHWC whatever ``data_format`` says; a symbolic tensor only REPORTS its
shape in the layout the reference asked for.  A sequence is a list of per-frame nodes of batch B; merging / splitting the
(B, T) axes (tf.reshape in DownBlock2D.call / ULSTMnet2D.call) is bookkeeping.
"""
import sys
import types

import numpy as np

from tests import tf_graphdef as G
from tests.keras_standin import _PermissiveModule, load_weights  # noqa: F401  (load_weights is re-exported)


class Graph:
    """One GraphDef under construction: a single placeholder holding all frames of all calls along the channel axis."""

    def __init__(self, B, H, W, C, T, n_calls, channels_first):
        self.B, self.H, self.W, self.C, self.T, self.n_calls = B, H, W, C, T, n_calls
        self.channels_first = channels_first
        self.c_total = n_calls * T * C
        self.bytes = G.placeholder('inp', [B, H, W, self.c_total])
        self.fetch = {}                     # node -> name to fetch it by (OpenCV folds BiasAdd into the Conv2D layer)
        self.n = 0

    def uid(self, base):
        self.n += 1
        return '%s_%d' % (base, self.n)

    def add(self, b):
        self.bytes += b

    def input(self, call):
        frames = []
        for t in range(self.T):
            nm = 'call%d/frame_%d' % (call, t)
            self.add(G.select_channels(nm, 'inp', self.c_total, (call * self.T + t) * self.C, self.C))
            frames.append(nm)
        return Sym(self, frames, self.B, self.H, self.W, self.C, merged=False)

    def pack_input(self, xs):
        """xs: list over calls of (B, T, C, H, W) / (B, T, H, W, C) arrays in the model's data_format -> NHWC placeholder value"""
        parts = []
        for x in xs:
            x = np.asarray(x, np.float32)
            if self.channels_first:
                x = x.transpose(0, 1, 3, 4, 2)
            parts += [x[:, t] for t in range(self.T)]
        return np.concatenate(parts, axis=-1)


class Sym:
    def __init__(self, g, frames, B, H, W, C, merged, crop=None):
        self.g, self.frames, self.B, self.H, self.W, self.C, self.merged, self.crop = g, list(frames), B, H, W, C, merged, crop

    @property
    def T(self):
        return len(self.frames)

    @property
    def shape(self):
        sp = (self.C, self.H, self.W) if self.g.channels_first else (self.H, self.W, self.C)
        return ((self.B * self.T,) if self.merged else (self.B, self.T)) + sp

    def like(self, frames, H=None, W=None, C=None, merged=None):
        return Sym(self.g, frames, self.B, self.H if H is None else H, self.W if W is None else W, self.C if C is None else C,
                   self.merged if merged is None else merged, self.crop)

    def __getitem__(self, idx):              # the crop of ULSTMnet2D.call (Networks.py:250): full batch / time / channel ranges
        assert not self.merged and len(idx) == 5 and self.crop is None
        sl = [(int(s.start or 0), int(s.stop)) for s in idx]
        cf = self.g.channels_first
        (b0, b1), (t0, t1) = sl[0], sl[1]
        (c0, c1), (y0, y1), (x0, x1) = (sl[2], sl[3], sl[4]) if cf else (sl[4], sl[2], sl[3])
        assert (b0, b1) == (0, self.B) and (t0, t1) == (0, self.T) and (c0, c1) == (0, self.C), sl
        assert 0 <= y0 < y1 <= self.H and 0 <= x0 < x1 <= self.W
        out = self.like(self.frames, H=y1 - y0, W=x1 - x0)
        out.crop = (y0, y1, x0, x1)
        return out


# ---- tf functions ---------------------------------------------------------------------------------------------------------
def _mod(a, b):
    return int(a) % int(b)


def _reshape(x, shape):
    shape = [int(s) for s in shape]
    if len(shape) == 4:
        assert not x.merged and shape == [x.B * x.T] + list(x.shape[2:]), (shape, x.shape)
        return x.like(x.frames, merged=True)
    assert x.merged and len(shape) == 5 and shape[0] * shape[1] == x.B * x.T and shape[0] == x.B and shape[2:] == list(x.shape[1:]), \
        (shape, x.shape)
    return x.like(x.frames, merged=False)


def _pad(x, paddings, mode):
    assert mode == 'REFLECT' and not x.merged and len(paddings) == 5
    p = [[int(a), int(b)] for a, b in paddings]
    cf = x.g.channels_first
    py, px = (p[3], p[4]) if cf else (p[2], p[3])
    rest = [p[0], p[1], p[2] if cf else p[4]]
    assert all(q == [0, 0] for q in rest)
    if not (max(py) or max(px)):
        return x
    frames = []
    for f in x.frames:
        nm = x.g.uid('pad')
        x.g.add(G.mirror_pad(nm, f, py[0], py[1], px[0], px[1]))
        frames.append(nm)
    return x.like(frames, H=x.H + sum(py), W=x.W + sum(px))


def _concat(xs, axis):
    a, b = xs
    assert a.merged and b.merged and axis in (1, -1) and (axis == 1) == a.g.channels_first
    assert (a.B, a.T, a.H, a.W) == (b.B, b.T, b.H, b.W)
    frames = []
    for fa, fb in zip(a.frames, b.frames):
        nm = a.g.uid('concat')
        a.g.add(G.concat(nm, [fa, fb]))
        frames.append(nm)
    return a.like(frames, C=a.C + b.C)


def _resize_images(x, hf, wf, data_format, interpolation='nearest'):
    assert interpolation == 'bilinear' and x.merged and (data_format == 'channels_first') == x.g.channels_first
    frames = []
    for f in x.frames:
        nm = x.g.uid('resize')
        x.g.add(G.resize_bilinear(nm, f, x.H * hf, x.W * wf))
        frames.append(nm)
    return x.like(frames, H=x.H * hf, W=x.W * wf)


# ---- keras layers -----------------------------------------------------------------------------------------------------------
def _np(a):
    return np.asarray(a.detach().numpy() if hasattr(a, 'detach') else a, dtype=np.float32)


class Model:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


class ConvLSTM2D:
    def __init__(self, filters, kernel_size, strides, padding, data_format, return_sequences, stateful):
        assert strides == 1 and padding == 'same' and return_sequences and stateful
        self.filters, self.kernel_size, self.data_format = filters, kernel_size, data_format
        self.state = None                    # (h node, c node) after the last call: Keras stateful=True

    def __call__(self, x):
        assert not x.merged and (self.data_format == 'channels_first') == x.g.channels_first
        g = x.g
        wk, wr, b = _np(self.kernel), _np(self.recurrent_kernel), _np(self.bias)
        assert wk.shape == (self.kernel_size, self.kernel_size, x.C, 4 * self.filters)
        if self.state is None:               # first call: zero states of the input's batch and size
            h, c = g.uid('h_init'), g.uid('c_init')
            g.add(G.zeros_like_channels(h, x.frames[0], x.C, self.filters) + G.zeros_like_channels(c, x.frames[0], x.C, self.filters))
        else:
            h, c = self.state
        outs = []
        for f in x.frames:
            gg, h, c = G.convlstm_cell(g.uid('convlstm'), f, h, c, wk, wr, b)
            g.add(gg)
            outs.append(h)
        self.state = (h, c)
        return x.like(outs, C=self.filters)


class Conv2D:
    def __init__(self, filters, kernel_size, strides, use_bias, data_format, padding):
        assert padding == 'same' and use_bias
        self.filters, self.kernel_size, self.strides, self.data_format = filters, kernel_size, strides, data_format

    def __call__(self, x):
        assert x.merged and (self.data_format == 'channels_first') == x.g.channels_first
        w, b = _np(self.kernel), _np(self.bias)
        assert w.shape == (self.kernel_size, self.kernel_size, x.C, self.filters)
        frames = []
        for f in x.frames:
            nm = x.g.uid('conv')
            x.g.add(G.conv2d(nm, f, w, self.strides, b))
            x.g.fetch[nm] = nm + '/Conv2D'
            frames.append(nm)
        s = self.strides
        return x.like(frames, H=-(-x.H // s), W=-(-x.W // s), C=self.filters)


class BatchNormalization:
    def __init__(self, axis):
        self.axis = axis

    def __call__(self, x, training=None):
        assert not training, 'inference graph: moving statistics'
        assert x.merged and (self.axis == 1) == x.g.channels_first
        frames = []
        for f in x.frames:
            nm = x.g.uid('bn')
            x.g.add(G.batchnorm(nm, f, _np(self.gamma), _np(self.beta), _np(self.moving_mean), _np(self.moving_variance)))
            frames.append(nm)
        return x.like(frames)


class LeakyReLU:
    def __call__(self, x):
        frames = []
        for f in x.frames:
            nm = x.g.uid('lrelu')
            x.g.add(G.leaky_relu(nm, f))
            frames.append(nm)
        return x.like(frames)


class Softmax:
    def __init__(self, axis=-1):
        self.axis = axis

    def __call__(self, x):
        assert not x.merged
        assert x.g.channels_first and self.axis == 2, 'channel soft-max of a (B, T, C, H, W) tensor (the channels-last quirk is not a graph op here)'
        frames = []
        for f in x.frames:
            nm = x.g.uid('softmax')
            x.g.add(G.softmax(nm, f))
            frames.append(nm)
        return x.like(frames)


def install():
    """Registers the stand-in as ``tensorflow`` / ``tensorflow.python.keras``; returns a function that removes it."""
    tf = _PermissiveModule('tensorflow')
    tf.__version__ = '2.0.graph-standin'
    tf.pad, tf.reshape, tf.concat = _pad, _reshape, _concat
    tf.math = types.SimpleNamespace(mod=_mod)
    keras = types.ModuleType('tensorflow.python.keras')
    keras.Model = Model
    keras.layers = types.SimpleNamespace(ConvLSTM2D=ConvLSTM2D, Conv2D=Conv2D, BatchNormalization=BatchNormalization,
                                         LeakyReLU=LeakyReLU, Softmax=Softmax)
    keras.backend = types.SimpleNamespace(resize_images=_resize_images)
    py = types.ModuleType('tensorflow.python')
    py.keras = keras
    tf.python = py
    names = {'tensorflow': tf, 'tensorflow.python': py, 'tensorflow.python.keras': keras}
    saved = {n: sys.modules.get(n) for n in names}
    sys.modules.update(names)

    def remove():
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
    return remove


def evaluate(g, xs, outputs):
    """Runs the graph with OpenCV on the packed inputs; outputs: list of Sym (un-merged) -> list of (B, T, C, H, W) arrays
    (channels_first) with each Sym's crop applied."""
    names, spans = [], []
    for s in outputs:
        assert not s.merged
        spans.append((len(names), len(names) + s.T))
        names += [g.fetch.get(f, f) for f in s.frames]
    uniq = list(dict.fromkeys(names))
    vals = dict(zip(uniq, G.run_with_opencv(g.bytes, g.pack_input(xs), uniq)))
    res = []
    for s, (a, b) in zip(outputs, spans):
        arr = np.stack([vals[n] for n in names[a:b]], axis=1)           # (B, T, C, H, W)
        if s.crop is not None:
            y0, y1, x0, x1 = s.crop
            arr = arr[:, :, :, y0:y1, x0:x1]
        res.append(arr if g.channels_first else arr.transpose(0, 1, 3, 4, 2))
    return res
