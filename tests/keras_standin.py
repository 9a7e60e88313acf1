"""TEST-ONLY stand-in for the few TensorFlow / Keras names the reference's Networks.py uses, backed by torch on the CPU.

Purpose: run the reference's OWN network code -- ``ULSTMnet2D.call`` with its padding / crop arithmetic, the block
wiring, the skip order, the reshape conventions, ``return_logits``, the soft-max axis, the state methods
(/root/reference/Networks.py, imported unmodified) -- on real numbers and compare it with ``oracle.OracleNet``.  The layer
ARITHMETIC inside the stand-in is the oracle's own statement of the Keras-2 operators (oracle/lstm_unet_oracle.py, SURVEY
App. A), so what this pins is everything in the oracle EXCEPT those operator semantics, which stay unpinned without a
TensorFlow installation.  Tensors are torch tensors in the layouts Keras would see (NCHW / NHWC as ``data_format`` says).
"""
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

from oracle import lstm_unet_oracle as O


def _to_nchw(x, channels_first):
    return x if channels_first else x.permute(0, 3, 1, 2)


def _from_nchw(x, channels_first):
    return x if channels_first else x.permute(0, 2, 3, 1)


class _Variable:
    """the ``.assign`` / ``.numpy`` surface of a tf.Variable around a torch tensor (API layout)"""

    def __init__(self, t):
        self.t = t

    def assign(self, v):
        self.t = v.t if isinstance(v, _Variable) else torch.as_tensor(v, dtype=self.t.dtype)

    def numpy(self):
        return self.t.numpy().copy()

    def __mul__(self, other):
        return self.t * other


class Model:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)

    def load_weights(self, prefix):
        """keras.Model.load_weights on a TF2 checkpoint prefix: through the repository's tensor-bundle reader"""
        from lstm_unet_b200 import tf_checkpoint
        names = [n for n, _, _ in O.build_param_specs(self._standin_net_params)]
        named = tf_checkpoint.load_model_weights(str(prefix), names)
        load_weights(self, {k: torch.from_numpy(np.array(v, dtype=np.float32)) for k, v in named.items()})


class ConvLSTM2D:
    def __init__(self, filters, kernel_size, strides, padding, data_format, return_sequences, stateful):
        assert strides == 1 and padding == 'same' and return_sequences and stateful
        self.filters, self.k, self.cf = filters, kernel_size, data_format == 'channels_first'
        self.kernel = self.recurrent_kernel = self.bias = None
        self.states = [None, None]

    def __call__(self, x5):
        x5 = x5 if self.cf else x5.permute(0, 1, 4, 2, 3)
        B, T = x5.shape[:2]
        if self.states[0] is None:
            h = torch.zeros((B, self.filters, x5.shape[3], x5.shape[4]), dtype=x5.dtype)
            c = torch.zeros_like(h)
        else:
            h, c = (_to_nchw(s.t, self.cf) for s in self.states)
        outs = []
        for t in range(T):
            z = O.conv2d_same(x5[:, t], self.kernel, self.bias, 1) + O.conv2d_same(h, self.recurrent_kernel, None, 1)
            zi, zf, zc, zo = torch.split(z, self.filters, dim=1)
            i, f, o = O.hard_sigmoid(zi), O.hard_sigmoid(zf), O.hard_sigmoid(zo)
            c = f * c + i * torch.tanh(zc)
            h = o * torch.tanh(c)
            outs.append(h)
        self.states = [_Variable(_from_nchw(h, self.cf).contiguous()), _Variable(_from_nchw(c, self.cf).contiguous())]
        y = torch.stack(outs, 1)
        return y if self.cf else y.permute(0, 1, 3, 4, 2)

    def reset_states(self, states=None):
        if states is None:
            self.states = [None, None]
        else:
            self.states = [_Variable(torch.as_tensor(np.asarray(s), dtype=torch.float32).clone()) for s in states]


class Conv2D:
    def __init__(self, filters, kernel_size, strides, use_bias, data_format, padding):
        assert padding == 'same' and use_bias
        self.stride, self.cf = strides, data_format == 'channels_first'
        self.kernel = self.bias = None

    def __call__(self, x):
        return _from_nchw(O.conv2d_same(_to_nchw(x, self.cf), self.kernel, self.bias, self.stride), self.cf)


class BatchNormalization:
    def __init__(self, axis):
        self.cf = axis == 1
        self.gamma = self.beta = self.moving_mean = self.moving_variance = None

    def __call__(self, x, training=None):
        y = O.batchnorm(_to_nchw(x, self.cf), self.gamma, self.beta, self.moving_mean, self.moving_variance, bool(training))
        return _from_nchw(y, self.cf)


class LeakyReLU:
    def __call__(self, x):
        return O.leaky_relu(x)


class Softmax:
    def __init__(self, axis=-1):
        self.axis = axis

    def __call__(self, x):
        return torch.softmax(x, dim=self.axis)


def _resize_images(x, hf, wf, data_format, interpolation='nearest'):
    assert hf == wf and interpolation == 'bilinear'
    cf = data_format == 'channels_first'
    return _from_nchw(O.resize_bilinear(_to_nchw(x, cf), hf), cf)


def _pad(x, paddings, mode):
    """tf.pad(x, paddings, 'REFLECT'): mirror without repeating the edge, dimension by dimension"""
    assert mode == 'REFLECT'
    for d, (lo, hi) in enumerate(paddings):
        lo, hi = int(lo), int(hi)
        if lo == 0 and hi == 0:
            continue
        n = x.shape[d]
        assert lo < n and hi < n
        idx = [lo - i for i in range(lo)] + list(range(n)) + [n - 2 - i for i in range(hi)]
        x = torch.index_select(x, d, torch.tensor(idx, dtype=torch.long))
    return x


class _Anything:
    """whatever else the reference touches at import time (default arguments such as tf.train.Coordinator())"""

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


class _PermissiveModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Anything()


def install():
    """Registers the stand-in as ``tensorflow`` / ``tensorflow.python.keras``; returns a function that removes it."""
    tf = _PermissiveModule('tensorflow')
    tf.__version__ = '2.0.standin'
    tf.pad = _pad
    tf.reshape = lambda x, shape: torch.as_tensor(x).reshape([int(s) for s in shape])
    tf.concat = lambda xs, axis: torch.cat(list(xs), dim=axis)
    tf.math = types.SimpleNamespace(mod=lambda a, b: int(a) % int(b))

    class _Device:                            # `with tf.device('/cpu:0'):`
        def __init__(self, name):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False
    tf.device = _Device
    # tf.data.Dataset.from_generator(gen, dtype): the generator itself (Inference2D.py iterates it)
    tf.data = types.SimpleNamespace(Dataset=types.SimpleNamespace(from_generator=lambda gen, dtype: gen()))
    # the tensor ops losses.WeightedCELoss uses (losses.py:13-27)
    tf.float32, tf.int32 = torch.float32, torch.int32
    tf.squeeze = lambda x, axis: torch.squeeze(torch.as_tensor(x), dim=axis)
    tf.cast = lambda x, dtype: torch.as_tensor(x).to(dtype)
    tf.greater = lambda a, b: torch.gt(torch.as_tensor(a), b)
    tf.transpose = lambda x, perm: torch.as_tensor(x).permute(*perm)
    tf.constant = lambda v: torch.tensor(v, dtype=torch.float32)
    tf.maximum = lambda a, b: torch.clamp(a, min=b)

    def one_hot(idx, depth):                  # out-of-range indices (the ignore label -1) give an all-zero row
        idx = idx.long()
        valid = (idx >= 0) & (idx < depth)
        return F.one_hot(torch.where(valid, idx, torch.zeros_like(idx)), depth).to(torch.float32) * valid.unsqueeze(-1)
    tf.one_hot = one_hot

    def reduce_sum(x, axis=None):
        return x.sum() if axis is None else x.sum(dim=axis)
    tf.reduce_sum = reduce_sum

    def sparse_ce(labels, logits):
        return torch.logsumexp(logits, dim=-1) - torch.gather(logits, -1, labels.long().unsqueeze(-1)).squeeze(-1)
    tf.nn = types.SimpleNamespace(sparse_softmax_cross_entropy_with_logits=sparse_ce)
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


def load_weights(model, params):
    """Gives the layers of a reference ``ULSTMnet2D`` the tensors of an oracle parameter dict (Keras variable names)."""
    for li, blk in enumerate(model.DownLayers):
        for j, l in enumerate(blk.ConvLSTM):
            pre = 'DownLayers/%d/ConvLSTM/%d/' % (li, j)
            l.kernel, l.recurrent_kernel, l.bias = params[pre + 'kernel'], params[pre + 'recurrent_kernel'], params[pre + 'bias']
        _load_convs(blk, 'DownLayers/%d/' % li, params)
    for ui, blk in enumerate(model.UpLayers):
        _load_convs(blk, 'UpLayers/%d/' % ui, params)


def _load_convs(blk, pre, params):
    for j, (c, bn) in enumerate(zip(blk.Conv, blk.BN)):
        c.kernel, c.bias = params[pre + 'Conv/%d/kernel' % j], params[pre + 'Conv/%d/bias' % j]
        if pre + 'BN/%d/gamma' % j in params:
            bn.gamma, bn.beta = params[pre + 'BN/%d/gamma' % j], params[pre + 'BN/%d/beta' % j]
            bn.moving_mean, bn.moving_variance = params[pre + 'BN/%d/moving_mean' % j], params[pre + 'BN/%d/moving_variance' % j]
