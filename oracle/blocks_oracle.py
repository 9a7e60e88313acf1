"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's blocks called ON THEIR OWN, the way
`DownBlock2D.unit_test` / `UpBlock2D.unit_test` construct and call them (Networks.py:100-119,155-175).

Only tests/ may import this.  Operator arithmetic is the network oracle's (oracle/lstm_unet_oracle.py: conv2d_same,
batchnorm, leaky_relu, resize_bilinear, hard_sigmoid -- SURVEY App. A); `tests/test_oracle.py` checks that these blocks
chained the way `ULSTMnet2D.call` chains them (Networks.py:233-245) reproduce `OracleNet`.  Like the network oracle the
operator semantics are **parity unpinned** against TensorFlow itself (not installable here); variable names follow the
network's (`DownLayers/0/...`, `UpLayers/0/...`) so that one parameter dictionary serves both sides of a comparison."""
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

from . import lstm_unet_oracle as O


def down_block_specs(conv_kernels, lstm_kernels, in_channels):
    net = {'down_conv_kernels': [list(conv_kernels)], 'lstm_kernels': [list(lstm_kernels)], 'up_conv_kernels': [[]]}
    return [s for s in O.build_param_specs(net, in_channels) if s[0].startswith('DownLayers/0/')]


def up_block_specs(kernels, in_channels, skip_channels, return_logits):
    specs, cin = [], in_channels + skip_channels
    for j, (k, f) in enumerate(kernels):
        specs += [('UpLayers/0/Conv/%d/kernel' % j, (k, k, cin, f), 'kernel'), ('UpLayers/0/Conv/%d/bias' % j, (f,), 'bias')]
        if not (return_logits and j == len(kernels) - 1):       # Networks.py:148-149: that BN is never called
            specs += [('UpLayers/0/BN/%d/gamma' % j, (f,), 'gamma'), ('UpLayers/0/BN/%d/beta' % j, (f,), 'beta'),
                      ('UpLayers/0/BN/%d/moving_mean' % j, (f,), 'moving_mean'),
                      ('UpLayers/0/BN/%d/moving_variance' % j, (f,), 'moving_var')]
        cin = f
    return specs


def init_from_specs(specs, seed=0, dtype=torch.float32):
    """Keras default initialisers with perturbed biases / BN variables (so that every term of the block is exercised)."""
    gen = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, shape, kind in specs:
        if kind == 'kernel':
            t = O._glorot_uniform(shape, gen, dtype)
        elif kind == 'recurrent':
            t = O._orthogonal(shape, gen, dtype)
        elif kind in ('gamma', 'moving_var'):
            t = (0.75 + 0.5 * torch.rand(shape, generator=gen, dtype=torch.float64)).to(dtype)
        else:
            t = 0.1 * torch.randn(shape, generator=gen, dtype=torch.float64).to(dtype)
            if kind == 'bias_lstm':
                f = shape[0] // 4
                t[f:2 * f] += 1.0                               # unit_forget_bias
        out[name] = t
    return out


class OracleDownBlock:
    """`DownBlock2D(conv_kernels, lstm_kernels, stride, data_format)` (Networks.py:37-98)."""

    def __init__(self, conv_kernels, lstm_kernels, stride=2, data_format='NCHW', params: Optional[Dict] = None,
                 in_channels=1, seed=0, dtype=torch.float32):
        self.conv_kernels, self.lstm_kernels, self.stride = list(conv_kernels), list(lstm_kernels), stride
        self.channels_first = data_format[1] == 'C'
        self.dtype = dtype
        self.params = params if params is not None else init_from_specs(
            down_block_specs(conv_kernels, lstm_kernels, in_channels), seed, dtype)
        self.states: List[Optional[List[torch.Tensor]]] = [None for _ in self.lstm_kernels]

    def __call__(self, inputs, training=False):
        x5 = torch.as_tensor(inputs, dtype=self.dtype)
        if not self.channels_first:
            x5 = x5.permute(0, 1, 4, 2, 3)
        p = self.params
        B, T = x5.shape[:2]
        for j in range(len(self.lstm_kernels)):                               # Networks.py:61-63
            pre = 'DownLayers/0/ConvLSTM/%d/' % j
            wk, wr, b = p[pre + 'kernel'], p[pre + 'recurrent_kernel'], p[pre + 'bias']
            Fo = wr.shape[2]
            if self.states[j] is None:
                h = torch.zeros((B, Fo, x5.shape[3], x5.shape[4]), dtype=self.dtype)
                c = torch.zeros_like(h)
            else:
                h, c = self.states[j]
            outs = []
            for t in range(T):
                z = O.conv2d_same(x5[:, t], wk, b, 1) + O.conv2d_same(h, wr, None, 1)
                zi, zf, zc, zo = torch.split(z, Fo, dim=1)
                c = O.hard_sigmoid(zf) * c + O.hard_sigmoid(zi) * torch.tanh(zc)
                h = O.hard_sigmoid(zo) * torch.tanh(c)
                outs.append(h)
            self.states[j] = [h.clone(), c.clone()]
            x5 = torch.stack(outs, dim=1)
        a = x5.reshape(B * T, *x5.shape[2:])                                    # Networks.py:65-67
        for j in range(len(self.conv_kernels)):                                 # Networks.py:69-72
            pre = 'DownLayers/0/'
            a = O.conv2d_same(a, p[pre + 'Conv/%d/kernel' % j], p[pre + 'Conv/%d/bias' % j], self.stride if j == 0 else 1)
            a = O.batchnorm(a, p[pre + 'BN/%d/gamma' % j], p[pre + 'BN/%d/beta' % j], p[pre + 'BN/%d/moving_mean' % j],
                            p[pre + 'BN/%d/moving_variance' % j], training)
            a = O.leaky_relu(a)
        down = a.reshape(B, T, *a.shape[1:])                                    # Networks.py:73-75
        if not self.channels_first:
            return down.permute(0, 1, 3, 4, 2), a.permute(0, 2, 3, 1)
        return down, a

    def reset_states_per_batch(self, is_last_batch):                            # Networks.py:77-84
        m = torch.as_tensor(is_last_batch, dtype=self.dtype).reshape(-1, 1, 1, 1)
        for st in self.states:
            if st is not None:
                st[0], st[1] = st[0] * m, st[1] * m


class OracleUpBlock:
    """`UpBlock2D(kernels, up_factor, data_format, return_logits)` (Networks.py:124-153)."""

    def __init__(self, kernels, up_factor=2, data_format='NCHW', return_logits=False, params: Optional[Dict] = None,
                 in_channels=1, skip_channels=1, seed=0, dtype=torch.float32):
        self.kernels, self.up_factor, self.return_logits = list(kernels), up_factor, return_logits
        self.channels_first = data_format[1] == 'C'
        self.dtype = dtype
        self.params = params if params is not None else init_from_specs(
            up_block_specs(kernels, in_channels, skip_channels, return_logits), seed, dtype)

    def __call__(self, inputs, training=False):
        x, skip = (torch.as_tensor(t, dtype=self.dtype) for t in inputs)
        if not self.channels_first:
            x, skip = x.permute(0, 3, 1, 2), skip.permute(0, 3, 1, 2)
        p = self.params
        x = torch.cat([O.resize_bilinear(x, self.up_factor), skip], dim=1)      # Networks.py:143-145
        for j in range(len(self.kernels)):
            x = O.conv2d_same(x, p['UpLayers/0/Conv/%d/kernel' % j], p['UpLayers/0/Conv/%d/bias' % j], 1)
            if self.return_logits and j == len(self.kernels) - 1:               # Networks.py:148-149
                break
            x = O.batchnorm(x, p['UpLayers/0/BN/%d/gamma' % j], p['UpLayers/0/BN/%d/beta' % j],
                            p['UpLayers/0/BN/%d/moving_mean' % j], p['UpLayers/0/BN/%d/moving_variance' % j], training)
            x = O.leaky_relu(x)
        return x if self.channels_first else x.permute(0, 2, 3, 1)
