"""CPU oracle for the ConvLSTM-UNet hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
leg may import this module.  The product path (``lstm_unet_b200``) never imports it and fails loudly
if its CUDA library is missing.

PARITY UNPINNED (operator semantics): the reference (arbellea/LSTM-UNet) delegates all arithmetic to
TensorFlow 2 / Keras-2 (``requirements.txt:1`` ``tensorflow_gpu>=2.0.0a0``, un-vendored, unpinned) and ships
no golden vectors or numeric tests.  TensorFlow is not installable in this image,
so the Keras-2 OPERATORS restated here (ConvLSTM2D, Conv2D SAME, BatchNormalization, LeakyReLU, bilinear
resize) cannot be checked against outputs of TensorFlow itself.  They are cross-checked against an independent
loop-level numpy restatement of the published semantics (``oracle/np_semantics.py``, tests/test_oracle.py).
and against a THIRD-PARTY implementation of the TensorFlow operators: tests/tf_graphdef.py writes the network's
forward as the GraphDef of TF ops Keras-2 lowers it to and OpenCV's TensorFlow importer executes it
(tests/test_tf_graph_opencv.py: operators, ConvLSTM steps and whole networks agree to 3e-6 relative).  That
is still not an output of the reference itself, hence the heading.

PINNED (wiring): everything around those operators IS checked against the reference's own code --
tests/test_reference_wiring.py imports /root/reference/Networks.py and losses.py unmodified on a torch-backed
stand-in for the handful of tf / Keras names they use (tests/keras_standin.py, whose layer arithmetic is this
module's) and compares ``ULSTMnet2D.call`` (reflect-pad / crop arithmetic, block and skip order, reshapes,
return_logits, the soft-max axis incl. the channels-last quirk, stateful carry), the state methods and
``WeightedCELoss`` with OracleNet / weighted_ce_loss on real numbers; tests/test_host_api.py pins the layer
hyper-parameters the reference's constructors pass to Keras and the Params defaults the same way.

What is restated (reference file:line -> function here):
  Networks.py:35-75    DownBlock2D.__init__/call          -> OracleNet._down_block
  Networks.py:77-98    per-sample state mask / get / set   -> reset_states_per_batch/get_states/set_states
  Networks.py:122-153  UpBlock2D                            -> OracleNet._up_block
  Networks.py:178-254  ULSTMnet2D.__init__/call             -> OracleNet.__init__/forward
  losses.py:8-27       WeightedCELoss                       -> weighted_ce_loss
  train2D.py:61,87-93  Adam + train_step                    -> keras_adam_step / train_step
Keras-2 layer semantics relied on implicitly by those lines:
  ConvLSTM2D (gate order i,f,c,o; hard_sigmoid recurrent activation; tanh; stateful; return_sequences),
  Conv2D SAME (TF asymmetric padding for stride 2), BatchNormalization (eps 1e-3, momentum .99, fused-BN
  unbiased moving variance), LeakyReLU(alpha=.3), resize_images bilinear half-pixel, tf.pad REFLECT.

All maths is done channels-first (NCHW) internally with torch CPU ops; weights are kept in the Keras
layouts (HWIO kernels; ConvLSTM kernels (k,k,Cin,4F)).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3          # keras BatchNormalization default epsilon
BN_MOMENTUM = 0.99     # keras BatchNormalization default momentum
LRELU_ALPHA = 0.3      # keras-2 LeakyReLU() default alpha

# Params.py:49-69 (CTCParams.net_kernel_params) -- the architecture train2D.py actually trains.
CTC_NET_PARAMS = {
    'down_conv_kernels': [[(3, 128), (3, 128)], [(3, 256), (3, 256)], [(3, 256), (3, 256)], [(3, 512), (3, 512)]],
    'lstm_kernels': [[(5, 128)], [(5, 256)], [(5, 256)], [(5, 512)]],
    'up_conv_kernels': [[(3, 256), (3, 256)], [(3, 128), (3, 128)], [(3, 64), (3, 64)], [(3, 32), (3, 32), (1, 3)]],
}


def hard_sigmoid(x: torch.Tensor) -> torch.Tensor:
    """Keras-2 ``hard_sigmoid``: clip(0.2*x + 0.5, 0, 1)  (ConvLSTM2D default recurrent_activation)."""
    return torch.clamp(0.2 * x + 0.5, 0.0, 1.0)


def tf_same_pad(in_size: int, k: int, s: int) -> Tuple[int, int]:
    """TensorFlow 'SAME' padding amounts (before, after) for one spatial dim."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    before = total // 2
    return before, total - before


def conv2d_same(x: torch.Tensor, w_hwio: torch.Tensor, b: Optional[torch.Tensor], stride: int) -> torch.Tensor:
    """Keras Conv2D(padding='same'): cross-correlation, HWIO kernel, TF SAME padding.  x is NCHW."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    pt, pb = tf_same_pad(x.shape[2], kh, stride)
    pl, pr = tf_same_pad(x.shape[3], kw, stride)
    x = F.pad(x, (pl, pr, pt, pb))
    return F.conv2d(x, w_hwio.permute(3, 2, 0, 1), b, stride=stride)


def batchnorm(x, gamma, beta, mov_mean, mov_var, training: bool, update: bool = True):
    """Keras BatchNormalization(axis=C) on NCHW.  Returns y; mutates moving stats in training."""
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if update:
            n = x.shape[0] * x.shape[2] * x.shape[3]
            with torch.no_grad():
                unbiased = var * (n / max(n - 1, 1))
                mov_mean.mul_(BN_MOMENTUM).add_((1 - BN_MOMENTUM) * mean)
                mov_var.mul_(BN_MOMENTUM).add_((1 - BN_MOMENTUM) * unbiased)
    else:
        mean, var = mov_mean, mov_var
    inv = torch.rsqrt(var + BN_EPS)
    return (x - mean[None, :, None, None]) * (inv * gamma)[None, :, None, None] + beta[None, :, None, None]


def leaky_relu(x):
    return torch.where(x > 0, x, LRELU_ALPHA * x)


def resize_bilinear(x: torch.Tensor, f: int) -> torch.Tensor:
    """k.backend.resize_images(..., interpolation='bilinear') == tf.image.resize half-pixel centres."""
    if f == 1:
        return x
    return F.interpolate(x, scale_factor=f, mode='bilinear', align_corners=False)


# ----------------------------------------------------------------------------------------------
# parameter construction (Keras default initialisers)
# ----------------------------------------------------------------------------------------------
def _glorot_uniform(shape, gen, dtype):
    rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    limit = math.sqrt(6.0 / (fan_in + fan_out))
    return ((torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) * limit).to(dtype)


def _orthogonal(shape, gen, dtype):
    rows, cols = int(np.prod(shape[:-1])), shape[-1]
    a = torch.randn((max(rows, cols), min(rows, cols)), generator=gen, dtype=torch.float64)
    q, r = torch.linalg.qr(a)
    q = q * torch.sign(torch.diagonal(r))[None, :]
    if rows < cols:
        q = q.t()
    return q[:rows, :cols].reshape(shape).to(dtype)


def build_param_specs(net_params: dict, in_channels: int = 1) -> List[Tuple[str, Tuple[int, ...], str]]:
    """Ordered list of (name, shape, kind) for every variable of ULSTMnet2D (Networks.py:179-206).

    kind in {'kernel','recurrent','bias_lstm','bias','gamma','beta','moving_mean','moving_var'}.
    The last conv of the last UpBlock returns logits before BN (Networks.py:148-149): no BN variables.
    """
    specs = []
    cin = in_channels
    skip_ch = []
    n_levels = len(net_params['down_conv_kernels'])
    for li in range(n_levels):
        skip_ch.append(cin)
        for j, (k, f) in enumerate(net_params['lstm_kernels'][li]):
            p = 'DownLayers/%d/ConvLSTM/%d/' % (li, j)
            specs += [(p + 'kernel', (k, k, cin, 4 * f), 'kernel'),
                      (p + 'recurrent_kernel', (k, k, f, 4 * f), 'recurrent'),
                      (p + 'bias', (4 * f,), 'bias_lstm')]
            cin = f
        for j, (k, f) in enumerate(net_params['down_conv_kernels'][li]):
            p = 'DownLayers/%d/' % li
            specs += [(p + 'Conv/%d/kernel' % j, (k, k, cin, f), 'kernel'), (p + 'Conv/%d/bias' % j, (f,), 'bias'),
                      (p + 'BN/%d/gamma' % j, (f,), 'gamma'), (p + 'BN/%d/beta' % j, (f,), 'beta'),
                      (p + 'BN/%d/moving_mean' % j, (f,), 'moving_mean'),
                      (p + 'BN/%d/moving_variance' % j, (f,), 'moving_var')]
            cin = f
    skip_ch.reverse()
    n_up = len(net_params['up_conv_kernels'])
    for ui in range(n_up):
        cin = cin + skip_ch[ui]
        convs = net_params['up_conv_kernels'][ui]
        for j, (k, f) in enumerate(convs):
            p = 'UpLayers/%d/' % ui
            specs += [(p + 'Conv/%d/kernel' % j, (k, k, cin, f), 'kernel'), (p + 'Conv/%d/bias' % j, (f,), 'bias')]
            is_logits = (ui == n_up - 1) and (j == len(convs) - 1)
            if not is_logits:
                specs += [(p + 'BN/%d/gamma' % j, (f,), 'gamma'), (p + 'BN/%d/beta' % j, (f,), 'beta'),
                          (p + 'BN/%d/moving_mean' % j, (f,), 'moving_mean'),
                          (p + 'BN/%d/moving_variance' % j, (f,), 'moving_var')]
            cin = f
    return specs


TRAINABLE_KINDS = ('kernel', 'recurrent', 'bias_lstm', 'bias', 'gamma', 'beta')


def init_params(net_params: dict, seed: int = 0, dtype=torch.float32, in_channels: int = 1,
                randomize_bn: bool = False) -> "OrderedDict[str, torch.Tensor]":
    """Keras default initialisation.  ``randomize_bn`` perturbs gamma/beta/moving
    stats and biases (test-only) so that BN folding and bias paths are actually exercised."""
    gen = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, shape, kind in build_param_specs(net_params, in_channels):
        if kind == 'kernel':
            t = _glorot_uniform(shape, gen, dtype)
        elif kind == 'recurrent':
            t = _orthogonal(shape, gen, dtype)
        elif kind == 'bias_lstm':
            f = shape[0] // 4
            t = torch.zeros(shape, dtype=dtype)
            t[f:2 * f] = 1.0  # unit_forget_bias
            if randomize_bn:
                t += 0.1 * torch.randn(shape, generator=gen, dtype=torch.float64).to(dtype)
        elif kind in ('bias', 'beta', 'moving_mean'):
            t = torch.zeros(shape, dtype=dtype)
            if randomize_bn:
                t += 0.1 * torch.randn(shape, generator=gen, dtype=torch.float64).to(dtype)
        else:  # gamma, moving_var
            t = torch.ones(shape, dtype=dtype)
            if randomize_bn:
                t *= (0.75 + 0.5 * torch.rand(shape, generator=gen, dtype=torch.float64)).to(dtype)
        out[name] = t
    return out


# ----------------------------------------------------------------------------------------------
# the network
# ----------------------------------------------------------------------------------------------
class OracleNet:
    """CPU restatement of ``Networks.ULSTMnet2D`` (Networks.py:178-291)."""

    def __init__(self, net_params: dict = CTC_NET_PARAMS, data_format: str = 'NCHW', pad_image: bool = True,
                 params: Optional[Dict[str, torch.Tensor]] = None, dtype=torch.float32, seed: int = 0,
                 gate: str = 'hard_sigmoid', in_channels: int = 1):
        if len(net_params['down_conv_kernels']) != len(net_params['lstm_kernels']):
            raise ValueError('Number of layers in down path ({}) do not match number of LSTM layers ({})'.format(
                len(net_params['down_conv_kernels']), len(net_params['lstm_kernels'])))
        if len(net_params['down_conv_kernels']) != len(net_params['up_conv_kernels']):
            raise ValueError('Number of layers in down path ({}) do not match number of layers in up path ({})'.format(
                len(net_params['down_conv_kernels']), len(net_params['up_conv_kernels'])))
        self.net_params = net_params
        self.channels_first = data_format[1] == 'C'       # Networks.py:181-182 (so 'NWHC' == channels-last)
        self.pad_image = pad_image
        self.dtype = dtype
        self.gate = hard_sigmoid if gate == 'hard_sigmoid' else torch.sigmoid
        self.n_levels = len(net_params['down_conv_kernels'])
        self.total_stride = 2 ** (self.n_levels - 1)       # Networks.py:197-199
        self.last_depth = net_params['up_conv_kernels'][-1][-1][1]
        self.params = params if params is not None else init_params(net_params, seed, dtype, in_channels)
        # stateful ConvLSTM states: [level][layer] -> [h, c] (NCHW) or None before first call
        self.states: List[List[Optional[List[torch.Tensor]]]] = [
            [None for _ in net_params['lstm_kernels'][li]] for li in range(self.n_levels)]

    # ---- ConvLSTM2D (keras defaults) ---------------------------------------
    def _conv_lstm(self, x5: torch.Tensor, li: int, j: int) -> torch.Tensor:
        p = self.params
        pre = 'DownLayers/%d/ConvLSTM/%d/' % (li, j)
        wk, wr, b = p[pre + 'kernel'], p[pre + 'recurrent_kernel'], p[pre + 'bias']
        B, T = x5.shape[0], x5.shape[1]
        Fo = wr.shape[2]
        st = self.states[li][j]
        if st is None:
            h = torch.zeros((B, Fo, x5.shape[3], x5.shape[4]), dtype=x5.dtype)
            c = torch.zeros_like(h)
        else:
            h, c = st[0].detach(), st[1].detach()      # truncated BPTT: no gradient into earlier calls
            if h.shape[0] != B or h.shape[2:] != x5.shape[3:]:
                raise ValueError('stateful ConvLSTM state shape %s does not match input %s'
                                 % (tuple(h.shape), tuple(x5.shape)))
        outs = []
        for t in range(T):
            z = conv2d_same(x5[:, t], wk, b, 1) + conv2d_same(h, wr, None, 1)
            zi, zf, zc, zo = torch.split(z, Fo, dim=1)
            i, f, o = self.gate(zi), self.gate(zf), self.gate(zo)
            c = f * c + i * torch.tanh(zc)
            h = o * torch.tanh(c)
            outs.append(h)
        self.states[li][j] = [h.detach().clone(), c.detach().clone()]
        return torch.stack(outs, dim=1)

    # ---- DownBlock2D.call (Networks.py:60-75) -------------------------------------------------
    def _down_block(self, x5, li, training):
        for j in range(len(self.net_params['lstm_kernels'][li])):
            x5 = self._conv_lstm(x5, li, j)
        B, T = x5.shape[:2]
        a = x5.reshape(B * T, *x5.shape[2:])
        stride = 2 if li < self.n_levels - 1 else 1
        p = self.params
        for j in range(len(self.net_params['down_conv_kernels'][li])):
            pre = 'DownLayers/%d/' % li
            a = conv2d_same(a, p[pre + 'Conv/%d/kernel' % j], p[pre + 'Conv/%d/bias' % j], stride if j == 0 else 1)
            a = batchnorm(a, p[pre + 'BN/%d/gamma' % j], p[pre + 'BN/%d/beta' % j],
                          p[pre + 'BN/%d/moving_mean' % j], p[pre + 'BN/%d/moving_variance' % j], training)
            a = leaky_relu(a)
        return a.reshape(B, T, *a.shape[1:]), a

    # ---- UpBlock2D.call (Networks.py:141-153) -------------------------------------------------
    def _up_block(self, x, skip, ui, training):
        up_factor = 2 if ui > 0 else 1
        x = resize_bilinear(x, up_factor)
        x = torch.cat([x, skip], dim=1)
        convs = self.net_params['up_conv_kernels'][ui]
        p = self.params
        pre = 'UpLayers/%d/' % ui
        last_block = ui == len(self.net_params['up_conv_kernels']) - 1
        for j in range(len(convs)):
            x = conv2d_same(x, p[pre + 'Conv/%d/kernel' % j], p[pre + 'Conv/%d/bias' % j], 1)
            if last_block and j == len(convs) - 1:
                return x
            x = batchnorm(x, p[pre + 'BN/%d/gamma' % j], p[pre + 'BN/%d/beta' % j],
                          p[pre + 'BN/%d/moving_mean' % j], p[pre + 'BN/%d/moving_variance' % j], training)
            x = leaky_relu(x)
        return x

    # ---- ULSTMnet2D.call (Networks.py:208-254) ------------------------------------------------
    def forward(self, inputs, training: bool = False):
        x = torch.as_tensor(inputs, dtype=self.dtype)
        if x.dim() != 5:
            raise ValueError('expected a 5-D input (B,T,C,H,W) or (B,T,H,W,C)')
        if not self.channels_first:
            x = x.permute(0, 1, 4, 2, 3)
        B, T, C, H, W = x.shape
        s = self.total_stride
        min_pad = s if self.pad_image else 0
        pad_y = (min_pad, min_pad + (s - H % s) % s)
        pad_x = (min_pad, min_pad + (s - W % s) % s)
        if max(pad_y) >= H or max(pad_x) >= W:
            raise ValueError('REFLECT padding needs pad < dim')
        xp = F.pad(x.reshape(B * T, C, H, W), (pad_x[0], pad_x[1], pad_y[0], pad_y[1]), mode='reflect') \
            if (max(pad_y) or max(pad_x)) else x.reshape(B * T, C, H, W)
        Hp, Wp = xp.shape[2], xp.shape[3]
        out_down = xp.reshape(B, T, C, Hp, Wp)
        out_skip = xp
        skips = []
        for li in range(self.n_levels):
            skips.append(out_skip)                      # appended BEFORE the block (Networks.py:239)
            out_down, out_skip = self._down_block(out_down, li, training)
        up = out_skip
        skips.reverse()
        for ui in range(len(self.net_params['up_conv_kernels'])):
            up = self._up_block(up, skips[ui], ui, training)
        logits = up.reshape(B, T, *up.shape[1:])
        logits = logits[:, :, :self.last_depth, pad_y[0]:pad_y[0] + H, pad_x[0]:pad_x[0] + W]
        if self.channels_first:
            softmax = torch.softmax(logits, dim=2)       # Softmax(channel_axis + 1) with channel_axis = 1
        else:
            logits = logits.permute(0, 1, 3, 4, 2)
            softmax = torch.softmax(logits, dim=0)       # reference quirk: Softmax(-1 + 1) == batch axis
        return logits, softmax

    __call__ = forward

    # ---- state API (Networks.py:77-98, 279-291) -----------------------------------------------
    def reset_states_per_batch(self, is_last_batch):
        m = torch.as_tensor(np.asarray(is_last_batch), dtype=self.dtype).reshape(-1, 1, 1, 1)
        for lvl in self.states:
            for st in lvl:
                if st is not None:
                    st[0] = st[0] * m
                    st[1] = st[1] * m

    def _to_api(self, t):
        return t.numpy().copy() if self.channels_first else t.permute(0, 2, 3, 1).numpy().copy()

    def _from_api(self, a):
        t = torch.as_tensor(np.asarray(a), dtype=self.dtype)
        return t.clone() if self.channels_first else t.permute(0, 3, 1, 2).contiguous()

    def get_states(self):
        return [[[None, None] if st is None else [self._to_api(st[0]), self._to_api(st[1])] for st in lvl]
                for lvl in self.states]

    def set_states(self, states):
        for li, lvl in enumerate(states):
            for j, st in enumerate(lvl):
                if st is None or st[0] is None:
                    cur = self.states[li][j]
                    self.states[li][j] = None if cur is None else [torch.zeros_like(cur[0]), torch.zeros_like(cur[1])]
                else:
                    self.states[li][j] = [self._from_api(st[0]), self._from_api(st[1])]

    def trainable_names(self):
        kinds = {n: k for n, _, k in build_param_specs(self.net_params, self._in_channels())}
        return [n for n in self.params if kinds[n] in TRAINABLE_KINDS]

    def _in_channels(self):
        return self.params['DownLayers/0/ConvLSTM/0/kernel'].shape[2]


# ----------------------------------------------------------------------------------------------
# loss / optimiser / train step
# ----------------------------------------------------------------------------------------------
def weighted_ce_loss(labels, logits, class_weights: Sequence[float], channels_first: bool = True):
    """losses.py:13-27.  labels: (B,T,1,H,W) / (B,T,H,W,1) floats in {-1,0,1,2}; logits (B,T,3,H,W) / (...,3)."""
    lab = torch.as_tensor(labels)
    if channels_first:
        lab = lab.squeeze(2)
        lg = logits.permute(0, 1, 3, 4, 2)
    else:
        lab = lab.squeeze(-1)
        lg = logits
    valid = (lab > -1).to(lg.dtype)
    li = lab.to(torch.int64)
    cw = torch.as_tensor(class_weights, dtype=lg.dtype)
    onehot = F.one_hot(li.clamp(min=0), 3).to(lg.dtype) * (li >= 0).to(lg.dtype)[..., None]   # tf.one_hot(-1)=0
    pixel_w = (onehot * cw).sum(-1)
    li = li.clamp(min=0)
    ce = -torch.log_softmax(lg, dim=-1).gather(-1, li[..., None]).squeeze(-1)
    return (ce * pixel_w * valid).sum() / (valid.sum() + 0.00001)


def keras_adam_step(params: Dict[str, torch.Tensor], grads: Dict[str, torch.Tensor], m, v, step: int,
                    lr: float = 1e-5, b1: float = 0.9, b2: float = 0.999, eps: float = 1e-7):
    """Keras (TF2 OptimizerV2) Adam: eps outside the bias-corrected sqrt.  step is 1-based."""
    lr_t = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    with torch.no_grad():
        for n, g in grads.items():
            m[n].mul_(b1).add_((1 - b1) * g)
            v[n].mul_(b2).add_((1 - b2) * g * g)
            params[n].sub_(lr_t * m[n] / (torch.sqrt(v[n]) + eps))


def train_step(net: OracleNet, image, label, class_weights, m, v, step: int, lr: float = 1e-5):
    """train2D.py:87-93: fwd(training=True) -> loss -> grads -> Adam.  Returns (loss, logits, softmax, grads)."""
    names = net.trainable_names()
    for n in names:
        net.params[n].requires_grad_(True)
        net.params[n].grad = None
    logits, softmax = net.forward(image, True)
    loss = weighted_ce_loss(label, logits, class_weights, net.channels_first)
    loss.backward()
    grads = {n: net.params[n].grad.detach().clone() for n in names}
    for n in names:
        net.params[n].requires_grad_(False)
        net.params[n].grad = None
    keras_adam_step(net.params, grads, m, v, step, lr)
    return loss.detach(), logits.detach(), softmax.detach(), grads
