"""Drop-in mirror of the reference's ``Networks.py`` object protocol for the ConvLSTM-UNet hot path.

Same names, constructor arguments, call signature and error behaviour as the reference
(``/root/reference/Networks.py``); all arithmetic runs in hand-written sm_100a kernels behind the C-ABI of
``liblstm_unet_b200.so`` (include/lstm_unet_b200.h).  PyTorch tensors are device containers only.

    reference                                   here
    ---------------------------------------------------------------------------------------------
    DEFAULT_NET_DOWN_PARAMS (Networks.py:12-32)  DEFAULT_NET_DOWN_PARAMS (same values)
    ULSTMnet2D(net_params, data_format,          ULSTMnet2D(net_params, data_format, pad_image, *, precision=...,
               pad_image)     (:179)                        engine=..., gate=...)
    model(inputs, training) -> (logits,          same; inputs: numpy / torch (host or cuda), 5-D; outputs are cuda
          softmax)            (:208-254)          tensors with a ``.numpy()`` like tf.Tensor
    reset_states_per_batch / get_states /        same semantics (mask 1 = keep, 0 = reset; nested
          set_states          (:77-98,279-291)    [block][layer][h, c] numpy lists, None before the first call)
    trainable_variables, save_weights,           named views into one flat fp32 device buffer; npz with Keras variable
          load_weights (keras.Model)              names and layouts
    DownBlock2D / UpBlock2D   (:35-175)           inside a network: descriptors of the blocks (state fan-out goes through
                                                  them); constructed on their own (the reference's unit_test usage): the same
                                                  kernels on a handle of their own (lu_config.block_kind)

There is no CPU fallback: constructing a model without the CUDA library or calling it without a CUDA device raises.
"""
import math

import numpy as np

from . import _lib
from .session import LuSession, LuError, TorchCudaBackend

__all__ = ['DEFAULT_NET_DOWN_PARAMS', 'DownBlock2D', 'UpBlock2D', 'ULSTMnet2D', 'Adam']

DEFAULT_NET_DOWN_PARAMS = {
    'down_conv_kernels': [
        [(5, 128), (5, 128)],
        [(5, 256), (5, 256)],
        [(5, 256), (5, 256)],
        [(5, 512), (5, 512)],
    ],
    'lstm_kernels': [
        [(5, 128)],
        [(5, 256)],
        [(5, 256)],
        [(5, 512)],
    ],
    'up_conv_kernels': [
        [(5, 256), (5, 256)],
        [(5, 128), (5, 128)],
        [(5, 64), (5, 64)],
        [(5, 32), (5, 32), (1, 3)],
    ],
}


_PINNED_POOL = {}
_POOL_KEEP = 4          # idle pinned buffers kept per (shape, dtype); more than that are released to the driver


def _pinned_for(t):
    """A pinned host buffer for the device -> host copy of `t` (DMA speed), owned by the caller through the array
    ``.numpy()`` returns: the buffer goes back to the pool only when that array (and every view of it) has been garbage
    collected, so results collected in a list are never overwritten -- like the fresh array tf.Tensor.numpy() returns
    (Inference2D.py:60) -- while a loop that drops each result reuses the same one or two buffers."""
    import torch
    import weakref
    key = (tuple(t.shape), t.dtype)
    pool = _PINNED_POOL.setdefault(key, [])
    free = [e for e in pool if e[1] is None or e[1]() is None]
    if free:
        entry = free[0]
        for extra in free[_POOL_KEEP:]:
            pool.remove(extra)
    else:
        entry = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True), None]
        pool.append(entry)
    arr = entry[0].numpy()
    entry[1] = weakref.ref(arr)
    return entry[0], arr


def _device_array_type():
    import torch

    class DeviceArray(torch.Tensor):
        """cuda tensor with tf.Tensor's ``.numpy()`` (device -> host copy), as Inference2D.py:60 uses it."""

        def numpy(self):
            t = self.as_subclass(torch.Tensor).detach()
            host, arr = _pinned_for(t)
            host.copy_(t, non_blocking=True)
            torch.cuda.current_stream(t.device).synchronize()
            return arr

    return DeviceArray


_DeviceArray = None


def _wrap(t):
    global _DeviceArray
    if _DeviceArray is None:
        _DeviceArray = _device_array_type()
    return t.as_subclass(_DeviceArray)


class Variable:
    """A named view into the flat parameter buffer (stands in for a tf.Variable of model.trainable_variables)."""

    def __init__(self, name, view, trainable, session=None):
        self.name, self.value, self.trainable = name, view, trainable
        self.shape = tuple(view.shape)
        self._session = session

    def numpy(self):
        return self.value.detach().cpu().numpy()

    def assign(self, arr):
        import torch
        self.value.copy_(torch.as_tensor(np.asarray(arr, dtype=np.float32)).reshape(self.value.shape))
        if self._session is not None:             # the library runs on packed bf16 copies / folded BN constants
            self._session.params_changed()


def _to_dev(x, dev):
    """numpy / torch (host or cuda) -> flat contiguous fp32 cuda tensor."""
    import torch
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
    return x.to(device=dev, dtype=torch.float32).contiguous().reshape(-1)


class _StandAloneBlock:
    """What a block needs to run outside a ULSTMnet2D: its own library handle (lu_config.block_kind), created by the
    first call like every Keras variable and state shape, default-initialised with the Keras initialisers."""
    _precision, _engine, _a_mode, _seed = 'bf16', 'tcgen05', 'halo', 0

    def _open(self, cfg):
        be = TorchCudaBackend(None)
        sess = LuSession(_lib.load_library(), be, cfg)
        sess.set_params(self._pending_weights or keras_default_init(sess.layout, self._seed))
        self._pending_weights = None
        self._be, self._sess = be, sess
        return sess

    def close(self):
        if getattr(self, '_sess', None) is not None:
            self._sess.close()
            self._sess = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weights_dict(self, named):
        """{variable name: array}; names as inside a network, block index 0 (`DownLayers/0/Conv/1/kernel`, ...)."""
        if getattr(self, '_sess', None) is None:
            self._pending_weights = {k: np.array(v, dtype=np.float32, copy=True) for k, v in named.items()}
        else:
            self._sess.set_params(named)

    def get_weights_dict(self):
        return self._sess.get_params()


class DownBlock2D(_StandAloneBlock):
    """One encoder block (Networks.py:35-98): ConvLSTM2D x n then [Conv2D, BN, LeakyReLU] x m.  Inside a ULSTMnet2D it
    is the structural descriptor the state fan-out goes through (the network's handle runs it); constructed on its own
    -- the reference's `DownBlock2D.unit_test` (Networks.py:100-119) -- it owns a handle of the same kernels and
    `block(inputs, training)` returns `(activ_down, activ)` like Networks.py:60-75."""

    def __init__(self, conv_kernels, lstm_kernels, stride=2, data_format='NCHW', _owner=None, _level=0, *,
                 precision='bf16', engine='tcgen05', a_mode='halo', seed=0):
        self.conv_kernels, self.lstm_kernels, self.stride = list(conv_kernels), list(lstm_kernels), stride
        self.data_format = data_format
        self.channel_axis = 1 if data_format[1] == 'C' else -1
        self.total_stride = stride                # Networks.py:52-54: only the first convolution is strided
        self._owner, self._level = _owner, _level
        self._precision, self._engine, self._a_mode, self._seed = precision, engine, a_mode, seed
        self._sess = self._pending_weights = self._shape = None

    # ---- on its own -------------------------------------------------------------------------------------
    def __call__(self, inputs, training=None, mask=None):
        import torch
        if self._owner is not None:
            raise NotImplementedError('this block belongs to a ULSTMnet2D: call the network')
        shape = tuple(inputs.shape)
        if len(shape) != 5:
            raise ValueError('expected a 5-D input, got shape %s' % (shape,))
        B, T = shape[:2]
        C, H, W = shape[2:] if self.channel_axis == 1 else (shape[4], shape[2], shape[3])
        if self._sess is not None and ((B, C, H, W) != self._shape or T > self._max_t):
            if (B, C, H, W) != self._shape:
                raise ValueError('stateful ConvLSTM states were built for (B,C,H,W)=%s, got %s' % (self._shape, (B, C, H, W)))
            states, weights = self.get_states(), self._sess.get_params()      # longer unroll: carry both over
            self.close()
            self._pending_weights = weights
            self._open_for(B, T, C, H, W)
            self.set_states(states)
        elif self._sess is None:
            self._open_for(B, T, C, H, W)
        sess, dev = self._sess, self._be.device
        n, F, Ho, Wo = sess.block_out_shape()
        out_shape = (B * T, F, Ho, Wo) if self.channel_axis == 1 else (B * T, Ho, Wo, F)
        out = torch.empty(out_shape, dtype=torch.float32, device=dev)
        xd = _to_dev(inputs, dev)
        sess.block_forward(xd.data_ptr(), None, T, bool(training), out.data_ptr())
        return _wrap(out.reshape((B, T) + out_shape[1:])), _wrap(out)       # Networks.py:73-75

    call = __call__

    def _open_for(self, B, T, C, H, W):
        try:
            self._open(_lib.make_down_block_config(self.conv_kernels, self.lstm_kernels, self.stride, self.data_format,
                                                   batch=B, max_t=T, height=H, width=W, in_channels=C,
                                                   precision=self._precision, engine=self._engine, a_mode=self._a_mode))
        except LuError as e:
            raise ValueError(str(e))
        self._shape, self._max_t = (B, C, H, W), T

    # ---- state API (Networks.py:77-98): through the network when the block is part of one -----------------
    def reset_states_per_batch(self, is_last_batch):
        import torch
        if self._owner is not None:
            return self._owner._reset_level(self._level, is_last_batch)
        if self._sess is None:
            return
        m = torch.as_tensor(np.asarray(is_last_batch, dtype=np.float32)).reshape(-1)
        if m.numel() != self._shape[0]:
            raise ValueError('mask has %d entries for batch size %d' % (m.numel(), self._shape[0]))
        md = m.to(self._be.device)
        self._sess.reset_level_states(0, md.data_ptr())
        torch.cuda.current_stream(self._be.device).synchronize()          # md must outlive the kernel

    def get_states(self):
        import torch
        if self._owner is not None:
            return self._owner._get_level_states(self._level)
        out = []
        for j in range(len(self.lstm_kernels)):
            if self._sess is None:
                out.append([None, None])
                continue
            shp = self._sess.state_shape(0, j)
            pair = []
            for which in (0, 1):
                t = torch.empty(shp, dtype=torch.float32, device=self._be.device)
                self._sess.get_state(0, j, which, t.data_ptr())
                pair.append(t.cpu().numpy())
            out.append(pair)
        return out

    def set_states(self, states):
        if self._owner is not None:
            return self._owner._set_level_states(self._level, states)
        if self._sess is None:
            raise ValueError('the states of a block exist after its first call')
        keep = []
        for j, st in enumerate(states):
            for which in (0, 1):
                if st is None or st[0] is None:                           # Networks.py:96-98: reset_states(None)
                    self._sess.set_state(0, j, which, None)
                else:
                    keep.append(_to_dev(st[which], self._be.device))
                    self._sess.set_state(0, j, which, keep[-1].data_ptr())
        self._be.synchronize()

    @classmethod
    def unit_test(cls):
        """The reference's shape check (Networks.py:100-119): 50 x 50 x 3 channels-last input, stride 2."""
        conv_kernels = [(3, 16), (3, 32), (3, 64)]
        lstm_kernels = [(3, 16), (3, 32), (3, 64)]
        model = cls(conv_kernels, lstm_kernels, 2, 'NHWC')
        for i in range(4):
            input_sequence = np.random.randn(2, 3, 50, 50, 3).astype(np.float32)
            model_out = model(input_sequence, True)
            print(i, tuple(model_out[0].shape), tuple(model_out[1].shape))
        return model_out


class UpBlock2D(_StandAloneBlock):
    """One decoder block (Networks.py:122-153): bilinear resize, concat([up, skip]), conv stack.  A descriptor inside a
    ULSTMnet2D; constructed on its own (`UpBlock2D.unit_test`, Networks.py:155-175) `block((inputs, skip), training)`
    runs it on a handle of its own."""

    def __init__(self, kernels, up_factor=2, data_format='NCHW', return_logits=False, *, precision='bf16',
                 engine='tcgen05', a_mode='halo', seed=0):
        self.kernels, self.up_factor, self.data_format, self.return_logits = list(kernels), up_factor, data_format, return_logits
        self.channel_axis = 1 if data_format[1] == 'C' else -1
        self._precision, self._engine, self._a_mode, self._seed = precision, engine, a_mode, seed
        self._sess = self._pending_weights = self._shape = None

    def __call__(self, inputs, training=None, mask=None):
        import torch
        x, skip = inputs
        xs, ss = tuple(x.shape), tuple(skip.shape)
        if len(xs) != 4 or len(ss) != 4:
            raise ValueError('expected 4-D (input, skip), got shapes %s and %s' % (xs, ss))
        N = xs[0]
        C, h, w = xs[1:] if self.channel_axis == 1 else (xs[3], xs[1], xs[2])
        Cs, H, W = ss[1:] if self.channel_axis == 1 else (ss[3], ss[1], ss[2])
        if ss[0] != N or (H, W) != (h * self.up_factor, w * self.up_factor):
            raise ValueError('skip of shape %s does not match the x%d up-sampled input of shape %s' % (ss, self.up_factor, xs))
        key = (N, C, h, w, Cs)
        if self._sess is not None and key != self._shape:
            weights = self._sess.get_params()                              # convolutions are shape-agnostic: new handle
            self.close()
            self._pending_weights = weights
        if self._sess is None:
            try:
                self._open(_lib.make_up_block_config(self.kernels, self.up_factor, self.data_format, self.return_logits,
                                                     frames=N, height=h, width=w, in_channels=C, skip_channels=Cs,
                                                     precision=self._precision, engine=self._engine, a_mode=self._a_mode))
            except LuError as e:
                raise ValueError(str(e))
            self._shape = key
        sess, dev = self._sess, self._be.device
        n, F, Ho, Wo = sess.block_out_shape()
        out = torch.empty((N, F, Ho, Wo) if self.channel_axis == 1 else (N, Ho, Wo, F), dtype=torch.float32, device=dev)
        xd, sd = _to_dev(x, dev), _to_dev(skip, dev)
        sess.block_forward(xd.data_ptr(), sd.data_ptr(), 1, bool(training), out.data_ptr())
        return _wrap(out)

    call = __call__

    @classmethod
    def unit_test(cls):
        """The reference's shape check (Networks.py:155-175)."""
        model = cls([(3, 16), (3, 32), (3, 64)], 2, 'NHWC')
        for i in range(4):
            input_sequence = np.random.randn(6, 50, 50, 3).astype(np.float32)
            skip = np.random.randn(6, 100, 100, 3).astype(np.float32)
            model_out = model((input_sequence, skip), True)
            print(i, tuple(model_out.shape))
        return model_out


def _glorot_uniform(shape, rng):
    rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    limit = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
    return rng.uniform(-limit, limit, size=shape).astype(np.float32)


def _orthogonal(shape, rng):
    rows, cols = int(np.prod(shape[:-1])), shape[-1]
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))[None, :]
    if rows < cols:
        q = q.T
    return q[:rows, :cols].reshape(shape).astype(np.float32)


def keras_default_init(layout, seed=0):
    """Keras default initialisers for every variable of the layout."""
    rng = np.random.default_rng(seed)
    out = {}
    for e in layout:
        name, shape = e['name'], e['shape']
        if name.endswith('recurrent_kernel'):
            out[name] = _orthogonal(shape, rng)
        elif name.endswith('kernel'):
            out[name] = _glorot_uniform(shape, rng)
        elif 'ConvLSTM' in name and name.endswith('bias'):
            b = np.zeros(shape, np.float32)
            f = shape[0] // 4
            b[f:2 * f] = 1.0                      # unit_forget_bias
            out[name] = b
        elif name.endswith('gamma') or name.endswith('moving_variance'):
            out[name] = np.ones(shape, np.float32)
        else:
            out[name] = np.zeros(shape, np.float32)
    return out


class Adam:
    """tf.keras.optimizers.Adam(lr) as train2D.py:61 constructs it (beta1 .9, beta2 .999, epsilon 1e-7, no amsgrad)."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, learning_rate=None):
        self.lr = learning_rate if learning_rate is not None else lr
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.iterations = 0
        self._m = self._v = None

    def get_slots(self):
        """(iterations, m, v) with the moment vectors as flat host arrays (None before the first step)."""
        if self._m is None:
            return self.iterations, None, None
        return self.iterations, self._m.detach().cpu().numpy().copy(), self._v.detach().cpu().numpy().copy()

    def set_slots(self, iterations, m, v, device=None):
        import torch
        self.iterations = int(iterations)
        if m is None:
            self._m = self._v = None
            return
        dev = device if device is not None else (self._m.device if self._m is not None else 'cpu')
        self._m = torch.as_tensor(np.asarray(m, dtype=np.float32)).to(dev)
        self._v = torch.as_tensor(np.asarray(v, dtype=np.float32)).to(dev)

    def _apply(self, model, grads):
        import torch
        sess = model._need_session()
        if self._m is not None and self._m.device != grads.device:      # restored from a checkpoint on the host
            self._m, self._v = self._m.to(grads.device), self._v.to(grads.device)
        if self._m is None:
            self._m = torch.zeros(sess.n_trainable, dtype=torch.float32, device=grads.device)
            self._v = torch.zeros_like(self._m)
        self.iterations += 1
        sess.adam_step(grads.data_ptr(), self._m.data_ptr(), self._v.data_ptr(), self.lr, self.iterations,
                       self.beta_1, self.beta_2, self.epsilon)


class ULSTMnet2D:
    def __init__(self, net_params=DEFAULT_NET_DOWN_PARAMS, data_format='NCHW', pad_image=True, *, precision='bf16',
                 engine='tcgen05', gate='hard_sigmoid', a_mode='halo', train=False, seed=0, device=None,
                 cuda_graph='auto', sync_bn=False, lrelu_alpha=0.3):
        # Networks.py:188-193: same ValueErrors, raised before anything touches the device
        _lib.make_config(net_params, data_format, pad_image)
        self.net_params = net_params
        self.data_format = data_format
        self.data_format_keras = 'channels_first' if data_format[1] == 'C' else 'channels_last'
        self.channel_axis = 1 if data_format[1] == 'C' else -1
        self.pad_image = pad_image
        self.precision, self.engine, self.gate, self.a_mode, self.train_capable = precision, engine, gate, a_mode, train
        # 'auto': the inference forward of launch-bound shapes (B*T <= 2, Inference2D's per-frame call) is replayed as a
        # CUDA graph; True / False force it where the library allows it (B*T <= 8, not a training model)
        self.cuda_graph = cuda_graph
        self.graph_active = False
        # data-parallel training: BatchNorm statistics over the batch of ALL ranks (parallel.enable_sync_batchnorm)
        self.sync_bn = sync_bn
        self.lrelu_alpha = lrelu_alpha            # Keras LeakyReLU() default; another value only for smooth-network tests
        self.seed, self._device = seed, device
        n = len(net_params['down_conv_kernels'])
        self.DownLayers = [DownBlock2D(c, l, 2 if i < n - 1 else 1, data_format, self, i)
                           for i, (c, l) in enumerate(zip(net_params['down_conv_kernels'], net_params['lstm_kernels']))]
        self.UpLayers = [UpBlock2D(c, 2 if i > 0 else 1, data_format, i + 1 == n)
                         for i, c in enumerate(net_params['up_conv_kernels'])]
        self.total_stride = 2 ** (n - 1)
        self.last_depth = net_params['up_conv_kernels'][-1][-1][1]
        self._lib = _lib.load_library()           # fails loudly when the CUDA library is not built
        self._sess = None
        self._pending_weights = None
        self._pending_states = None
        self._shape = None                        # (B, H, W) frozen by the first call (stateful ConvLSTM)
        self._x_pin = self._x_dev = None

    def close(self):
        """Release the library handle and the device workspace."""
        if getattr(self, '_sess', None) is not None:
            self._sess.close()
            self._sess = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- session management --------------------------------------------------------------------------
    def _build(self, B, T, H, W):
        be = TorchCudaBackend(self._device)
        old = self._sess
        old_states = self.get_states() if old is not None else None
        self._be = be
        cfg = _lib.make_config(self.net_params, self.data_format, self.pad_image, batch=B, max_t=T, height=H, width=W,
                               precision=self.precision, engine=self.engine, gate=self.gate, a_mode=self.a_mode,
                               train=self.train_capable, lrelu_alpha=self.lrelu_alpha,
                               in_channels=getattr(self, '_in_channels', 1))
        sess = LuSession(self._lib, be, cfg)
        want_graph = (B * T <= 2) if self.cuda_graph == 'auto' else bool(self.cuda_graph)
        self.graph_active = sess.set_graph_mode(want_graph and not self.train_capable)
        if self.sync_bn:
            from .parallel import enable_sync_batchnorm
            enable_sync_batchnorm(sess)
        if old is not None:                       # longer unroll than before: carry weights and states over
            sess.params.copy_(old.params)
            sess.params_changed()
            self._sess = sess
            self.set_states(old_states)
            old.close()
        else:
            w = self._pending_weights or keras_default_init(sess.layout, self.seed)
            if self._pending_weights is not None:
                k0 = 'DownLayers/0/ConvLSTM/0/kernel'
                if k0 in w and np.asarray(w[k0]).shape[2] != self._in_channels:
                    raise ValueError('weights were made for %d image channels, input has %d' % (np.asarray(w[k0]).shape[2], self._in_channels))
            sess.set_params(w)
            self._pending_weights = None
            self._sess = sess
            if self._pending_states is not None:
                self.set_states(self._pending_states)
                self._pending_states = None
        self._shape = (B, H, W)
        self._max_t = T
        return sess

    def _ensure(self, B, T, H, W):
        if self._sess is None:
            return self._build(B, T, H, W)
        if (B, H, W) != self._shape:
            raise ValueError('stateful ConvLSTM states were built for (B,H,W)=%s, got %s' % (self._shape, (B, H, W)))
        if T > self._max_t:
            return self._build(B, T, H, W)
        return self._sess

    def _check_channels(self, C):
        """The number of image channels is frozen by the first call, like every Keras variable shape."""
        if self._sess is not None and C != self._in_channels:
            raise ValueError('the model was built for %d image channels, got %d' % (self._in_channels, C))
        self._in_channels = int(C)

    # ---- call (Networks.py:208-254) ------------------------------------------------------------------
    def __call__(self, inputs, training=None, mask=None):
        import torch
        x = inputs
        if not isinstance(x, torch.Tensor):
            x = torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))
        if x.dim() != 5:
            raise ValueError('expected a 5-D input, got shape %s' % (tuple(x.shape),))
        if self.channel_axis == 1:
            B, T, C, H, W = x.shape
        else:
            B, T, H, W, C = x.shape
        self._check_channels(C)
        sess = self._ensure(B, T, H, W)
        dev = self._be.device
        if x.device.type != 'cuda':
            # host input: pinned staging buffer + async H2D on the compute stream
            n = x.numel()
            if self._x_pin is None or self._x_pin.numel() < n:
                self._x_pin = torch.empty(n, dtype=torch.float32, pin_memory=True)
                self._x_dev = torch.empty(n, dtype=torch.float32, device=dev)
                self._x_event = torch.cuda.Event()
            else:
                self._x_event.synchronize()       # the previous call's H2D may still be queued behind its forward
            self._x_pin[:n].copy_(x.reshape(-1).to(torch.float32))
            xd = self._x_dev[:n]
            xd.copy_(self._x_pin[:n], non_blocking=True)
            self._x_event.record(torch.cuda.current_stream(dev))
        else:
            xd = x.to(device=dev, dtype=torch.float32).contiguous().reshape(-1)
        shape = (B, T, self.last_depth, H, W) if self.channel_axis == 1 else (B, T, H, W, self.last_depth)
        logits = torch.empty(shape, dtype=torch.float32, device=dev)
        softmax = torch.empty(shape, dtype=torch.float32, device=dev)
        sess.forward(xd.data_ptr(), T, bool(training), logits.data_ptr(), softmax.data_ptr())
        lg, sm = _wrap(logits), _wrap(softmax)
        lg._lu_model = sm._lu_model = self          # lets losses.WeightedCELoss find the resident logits
        return lg, sm

    call = __call__

    def predict_batches(self, batches, training=False):
        """Pipelined inference over an iterable of HOST batches (numpy, 5-D, same shape): yields the soft-max of every
        batch, in order, as a host array the caller owns (Inference2D.py:59-60: ``model(image, training=False)`` followed
        by ``image_softmax.numpy()``).  Two slots: while batch i is computed, batch i+1 is copied host -> device and the
        soft-max of batch i-1 device -> host, each on its own stream, so the copies cost no step time.  Recurrent states
        carry from batch to batch exactly as with successive calls."""
        import collections
        import torch
        pending = collections.deque()
        slots, streams = [], None
        for i, x in enumerate(batches):
            x = np.ascontiguousarray(np.asarray(x, dtype=np.float32))
            if x.ndim != 5:
                raise ValueError('expected a 5-D input, got shape %s' % (x.shape,))
            if self.channel_axis == 1:
                B, T, C, H, W = x.shape
            else:
                B, T, H, W, C = x.shape
            self._check_channels(C)
            sess = self._ensure(B, T, H, W)
            dev = self._be.device
            comp = torch.cuda.current_stream(dev)
            if streams is None:
                streams = (torch.cuda.Stream(dev), torch.cuda.Stream(dev))
            h2d, d2h = streams
            shape = (B, T, self.last_depth, H, W) if self.channel_axis == 1 else (B, T, H, W, self.last_depth)
            if len(slots) < 2:
                slots.append({'x_pin': torch.empty(x.size, dtype=torch.float32, pin_memory=True),
                              'x_dev': torch.empty(x.size, dtype=torch.float32, device=dev),
                              'lg': torch.empty(shape, dtype=torch.float32, device=dev),
                              'sm': torch.empty(shape, dtype=torch.float32, device=dev),
                              'h2d_done': None, 'x_free': None, 'd2h_done': None})
            if tuple(slots[0]['sm'].shape) != shape or slots[0]['x_pin'].numel() != x.size:
                raise ValueError('every batch of a pipelined run must have the same shape')
            s = slots[i % 2]
            if s['h2d_done'] is not None:
                s['h2d_done'].synchronize()                       # the pinned staging buffer of batch i-2 has been read
            s['x_pin'].copy_(torch.from_numpy(x.reshape(-1)))
            with torch.cuda.stream(h2d):
                if s['x_free'] is not None:
                    h2d.wait_event(s['x_free'])                   # forward i-2 has consumed this slot's device input
                s['x_dev'].copy_(s['x_pin'], non_blocking=True)
                s['h2d_done'] = torch.cuda.Event()
                s['h2d_done'].record(h2d)
            comp.wait_event(s['h2d_done'])
            if s['d2h_done'] is not None:
                comp.wait_event(s['d2h_done'])                    # the soft-max of batch i-2 has left this slot
            sess.forward(s['x_dev'].data_ptr(), T, bool(training), s['lg'].data_ptr(), s['sm'].data_ptr())
            s['x_free'] = torch.cuda.Event()
            s['x_free'].record(comp)
            with torch.cuda.stream(d2h):
                d2h.wait_event(s['x_free'])
                host, arr = _pinned_for(s['sm'])
                host.copy_(s['sm'], non_blocking=True)
                s['d2h_done'] = torch.cuda.Event()
                s['d2h_done'].record(d2h)
            pending.append((s['d2h_done'], arr))
            if len(pending) == 2:
                ev, out = pending.popleft()
                ev.synchronize()
                yield out
        while pending:
            ev, out = pending.popleft()
            ev.synchronize()
            yield out

    # ---- recurrent state API (Networks.py:77-98, 279-291) ---------------------------------------------
    def _n_lstm(self, level):
        return len(self.net_params['lstm_kernels'][level])

    def _mask_dev(self, is_last_batch):
        import torch
        m = torch.as_tensor(np.asarray(is_last_batch, dtype=np.float32)).reshape(-1)
        if m.numel() != self._shape[0]:
            raise ValueError('mask has %d entries for batch size %d' % (m.numel(), self._shape[0]))
        return m.to(self._be.device)

    def _reset_level(self, level, is_last_batch):
        """DownBlock2D.reset_states_per_batch (Networks.py:77-84): the ConvLSTM layers of that block only."""
        import torch
        if self._sess is None:
            return
        md = self._mask_dev(is_last_batch)
        self._sess.reset_level_states(level, md.data_ptr())
        torch.cuda.current_stream(self._be.device).synchronize()      # md must outlive the kernel

    def reset_states_per_batch(self, is_last_batch):
        """ULSTMnet2D.reset_states_per_batch (Networks.py:279-281): every block."""
        import torch
        if self._sess is None:
            return
        md = self._mask_dev(is_last_batch)
        self._sess.reset_states(md.data_ptr())
        torch.cuda.current_stream(self._be.device).synchronize()      # md must outlive the kernel

    def _get_level_states(self, level):
        import torch
        out = []
        for j in range(self._n_lstm(level)):
            if self._sess is None:
                out.append([None, None])
                continue
            shp = self._sess.state_shape(level, j)
            pair = []
            for which in (0, 1):
                t = torch.empty(shp, dtype=torch.float32, device=self._be.device)
                self._sess.get_state(level, j, which, t.data_ptr())
                pair.append(t.cpu().numpy())
            out.append(pair)
        return out

    def get_states(self):
        return [self._get_level_states(i) for i in range(len(self.DownLayers))]

    def _set_level_states(self, level, states):
        import torch
        if self._sess is None:
            if self._pending_states is None:
                self._pending_states = [[[None, None] for _ in range(self._n_lstm(i))] for i in range(len(self.DownLayers))]
            self._pending_states[level] = states
            return
        for j, st in enumerate(states):
            shp = self._sess.state_shape(level, j)
            for which in (0, 1):
                if st is None or st[0] is None:
                    self._sess.set_state(level, j, which, None)     # keras reset_states(None): zeros
                else:
                    a = np.asarray(st[which], dtype=np.float32)
                    if tuple(a.shape) != shp:
                        raise ValueError('state shape %s does not match %s' % (a.shape, shp))
                    t = torch.from_numpy(np.ascontiguousarray(a)).to(self._be.device)
                    self._sess.set_state(level, j, which, t.data_ptr())
                    torch.cuda.current_stream(self._be.device).synchronize()

    def set_states(self, states):
        for i, st in enumerate(states):
            self._set_level_states(i, st)

    # ---- variables / weights (keras.Model surface used by train2D.py:92-93,236 and Inference2D.py:34) --
    def _need_session(self):
        if self._sess is None:
            raise RuntimeError('the model is built by its first call (Keras builds variables lazily too)')
        return self._sess

    @property
    def variables(self):
        s = self._need_session()
        return [Variable(e['name'], s.params[e['offset']:e['offset'] + e['count']].view(e['shape']), e['trainable'], s)
                for e in s.layout]

    @property
    def trainable_variables(self):
        return [v for v in self.variables if v.trainable]

    def get_weights_dict(self):
        return self._need_session().get_params()

    def set_weights_dict(self, named):
        if self._sess is None:
            self._pending_weights = {k: np.array(v, dtype=np.float32, copy=True) for k, v in named.items()}
        else:
            self._sess.set_params(named)

    def save_weights(self, path, save_format=None):
        """``save_format='tf'`` (train2D.py:236): a TF2 tensor bundle ``<path>.index`` + ``<path>.data-00000-of-00001`` with
        the Keras object-graph variable keys (written without TensorFlow, see tf_checkpoint.py).  Otherwise an ``.npz``
        with the Keras variable names / layouts."""
        w = self.get_weights_dict() if self._sess is not None else self._pending_weights
        if w is None:
            raise RuntimeError('nothing to save: the model has no weights yet')
        if save_format == 'tf':
            from . import tf_checkpoint
            tf_checkpoint.save_model_weights(str(path), w)
            return
        with open(path if str(path).endswith('.npz') else str(path) + '.npz', 'wb') as f:
            np.savez(f, **{k.replace('/', '|'): v for k, v in w.items()})

    def _variable_layout(self):
        """[{name, shape, offset, count, trainable}] of the flat parameter buffer (host logic only: works without a GPU)."""
        if self._sess is not None:
            return [dict(e) for e in self._sess.layout]
        cfg = _lib.make_config(self.net_params, self.data_format, self.pad_image, batch=1, max_t=1, height=64, width=64)
        import ctypes
        h = ctypes.c_void_p()
        if self._lib.lu_create(ctypes.byref(cfg), ctypes.byref(h)) != 0:
            raise LuError(self._lib.lu_last_error().decode())
        nt = ctypes.c_int32()
        self._lib.lu_param_count(h, ctypes.byref(nt), None, None)
        buf, out = ctypes.create_string_buffer(256), []
        for i in range(nt.value):
            shp, rank, off, tr = (ctypes.c_int64 * 4)(), ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int32()
            self._lib.lu_param_info(h, i, buf, 256, shp, ctypes.byref(rank), ctypes.byref(off), ctypes.byref(tr))
            shape = tuple(int(shp[j]) for j in range(rank.value))
            out.append({'name': buf.value.decode(), 'shape': shape, 'offset': off.value, 'count': int(np.prod(shape)),
                        'trainable': bool(tr.value)})
        self._lib.lu_destroy(h)
        return out

    def _variable_names(self):
        return [e['name'] for e in self._variable_layout()]

    def load_weights(self, path):
        """Accepts a TF2 checkpoint prefix (``model.ckpt`` -> ``model.ckpt.index``; Inference2D.py:34) or an ``.npz``."""
        import os
        if os.path.exists(str(path) + '.index'):
            from . import tf_checkpoint
            self.set_weights_dict(tf_checkpoint.load_model_weights(str(path), self._variable_names()))
            return
        p = path if str(path).endswith('.npz') else str(path) + '.npz'
        with np.load(p) as z:
            self.set_weights_dict({k.replace('|', '/'): z[k] for k in z.files})

    # ---- training step (train2D.py:87-93: GradientTape forward, WeightedCELoss, tape.gradient, Adam) -----------
    def _label_dev(self, label):
        import torch
        lab = label
        if not isinstance(lab, torch.Tensor):
            lab = torch.from_numpy(np.ascontiguousarray(np.asarray(lab, dtype=np.float32)))
        return lab.to(device=self._be.device, dtype=torch.float32).contiguous()

    def loss(self, label, class_weights):
        """WeightedCELoss (losses.py:13-27) of the logits of the LAST call; returns a 0-d cuda tensor."""
        import torch
        sess = self._need_session()
        lab = self._label_dev(label)
        out = torch.zeros(1, dtype=torch.float32, device=self._be.device)
        sess.loss_backward(lab.data_ptr(), class_weights, out.data_ptr(), None)
        self._keep = lab
        return out[0]

    def backward(self, label, class_weights):
        """loss + flat gradient of the trainable variables for the last call (training=True). -> (loss, grads)"""
        import torch
        sess = self._need_session()
        if not self.train_capable:
            raise RuntimeError('construct the model with train=True to use the backward pass')
        lab = self._label_dev(label)
        if getattr(self, '_grads', None) is None:
            self._grads = torch.zeros(sess.n_trainable, dtype=torch.float32, device=self._be.device)
        out = torch.zeros(1, dtype=torch.float32, device=self._be.device)
        sess.loss_backward(lab.data_ptr(), class_weights, out.data_ptr(), self._grads.data_ptr())
        self._keep = lab
        return out[0], self._grads

    def apply_gradients(self, grads, optimizer):
        """optimizer.apply_gradients(zip(grads, model.trainable_variables)) with Keras Adam (train2D.py:61,93)."""
        optimizer._apply(self, grads)

    def train_step(self, image, label, class_weights, optimizer, allreduce=None):
        """One train2D.train_step: returns (softmax, logits, loss).  `allreduce(flat_grads)` is the single data-parallel
        exchange (parallel.all_reduce_mean_)."""
        logits, softmax = self(image, True)
        if allreduce is not None and hasattr(allreduce, 'begin'):     # parallel.OverlappedAllReduce: buckets from the backward
            import torch
            if getattr(self, '_grads', None) is None:
                self._grads = torch.zeros(self._need_session().n_trainable, dtype=torch.float32, device=self._be.device)
            allreduce.begin(self._need_session(), self._grads)
        loss, grads = self.backward(label, class_weights)
        if allreduce is not None:
            allreduce(grads)
        optimizer._apply(self, grads)
        return softmax, logits, loss

    def forward_flops(self, T):
        return self._need_session().forward_flops(T)

    def launch_count(self, reset=False):
        return self._need_session().launch_count(reset)

    @classmethod
    def unit_test(cls):
        """Networks.py:256-277 (shape contract)."""
        model = cls(DEFAULT_NET_DOWN_PARAMS, 'NHWC', True)
        for i in range(4):
            x = np.random.randn(2, 2, 35, 35, 3).astype(np.float32)      # the reference's own case: 3 channels, channels-last
            out = model(x, True)
            print(i, tuple(out[0].shape))
