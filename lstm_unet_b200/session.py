"""LuSession: one library handle + its device buffers.  The buffer backend is injected: the product uses
TorchCudaBackend (torch tensors as containers only); the CPU test-suite injects a numpy backend together with the
TEST-ONLY host build of the kernels (tests/emu_backend.py).  Nothing in this package can run the path on a CPU."""
import ctypes

import numpy as np

from . import _lib


class LuError(ValueError):
    """Errors reported by the library (the reference raises ValueError for shape / configuration problems)."""


class TorchCudaBackend:
    name = 'torch-cuda'

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('lstm_unet_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
        self.torch = torch
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)

    def empty(self, n, dtype):
        t = self.torch
        td = {np.float32: t.float32, np.uint8: t.uint8, np.float64: t.float64}[dtype]
        return t.empty(int(n), dtype=td, device=self.device)

    def zeros(self, n, dtype):
        b = self.empty(n, dtype)
        b.zero_()
        return b

    def ptr(self, buf):
        return buf.data_ptr()

    def stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def to_device(self, arr, out=None):
        t = self.torch.from_numpy(np.ascontiguousarray(arr))
        if out is None:
            return t.to(self.device)
        out.copy_(t.reshape(out.shape), non_blocking=True)
        return out

    def to_host(self, buf):
        return buf.detach().cpu().numpy()

    def synchronize(self):
        self.torch.cuda.synchronize(self.device)


class LuSession:
    def __init__(self, lib, backend, cfg):
        self.lib, self.be, self.cfg = lib, backend, cfg
        h = ctypes.c_void_p()
        self._check(lib.lu_create(ctypes.byref(cfg), ctypes.byref(h)))
        self.h = h
        nb = ctypes.c_size_t()
        self._check(lib.lu_workspace_bytes(h, ctypes.byref(nb)))
        self.ws_bytes = nb.value
        self._ws_raw = backend.empty(nb.value + 1024, np.uint8)
        base = backend.ptr(self._ws_raw)
        self._ws_ptr = (base + 1023) // 1024 * 1024
        self._check(lib.lu_bind_workspace(h, self._ws_ptr, nb.value, backend.stream()))
        nt, ne, ntr = ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64()
        self._check(lib.lu_param_count(h, ctypes.byref(nt), ctypes.byref(ne), ctypes.byref(ntr)))
        self.n_elements, self.n_trainable = ne.value, ntr.value
        self.layout = []
        name = ctypes.create_string_buffer(256)
        for i in range(nt.value):
            shape = (ctypes.c_int64 * 4)()
            rank, off, tr = ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int32()
            self._check(lib.lu_param_info(h, i, name, 256, shape, ctypes.byref(rank), ctypes.byref(off), ctypes.byref(tr)))
            shp = tuple(int(shape[j]) for j in range(rank.value))
            self.layout.append({'name': name.value.decode(), 'shape': shp, 'offset': off.value,
                                'count': int(np.prod(shp)), 'trainable': bool(tr.value)})
        self.params = backend.zeros(self.n_elements, np.float32)
        self._check(lib.lu_bind_params(h, backend.ptr(self.params)))

    def _check(self, rc):
        if rc != 0:
            raise LuError(self.lib.lu_last_error().decode())

    def close(self):
        if self.h:
            self.lib.lu_destroy(self.h)
            self.h = None

    # ---- parameters -----------------------------------------------------------------------------------
    def set_params(self, named):
        """named: {keras name: array in keras layout}.  Missing names raise KeyError."""
        flat = np.empty(self.n_elements, dtype=np.float32)
        for e in self.layout:
            a = np.asarray(named[e['name']], dtype=np.float32)
            if tuple(a.shape) != e['shape']:
                raise LuError('parameter %s has shape %s, expected %s' % (e['name'], a.shape, e['shape']))
            flat[e['offset']:e['offset'] + e['count']] = a.reshape(-1)
        self.be.to_device(flat, out=self.params)
        self.params_changed()

    def get_params(self):
        flat = self.be.to_host(self.params)
        return {e['name']: flat[e['offset']:e['offset'] + e['count']].reshape(e['shape']).copy() for e in self.layout}

    def params_changed(self):
        self._check(self.lib.lu_params_changed(self.h, self.be.stream()))

    # ---- forward / state ------------------------------------------------------------------------------
    def forward(self, x_ptr, T, training, logits_ptr, softmax_ptr):
        self._check(self.lib.lu_forward(self.h, x_ptr, int(T), 1 if training else 0, logits_ptr, softmax_ptr,
                                        self.be.stream()))

    def block_forward(self, x_ptr, skip_ptr, T, training, out_ptr):
        """A stand-alone DownBlock2D / UpBlock2D handle (lu_block_forward); skip_ptr is None for a DownBlock2D."""
        self._check(self.lib.lu_block_forward(self.h, x_ptr, skip_ptr, int(T), 1 if training else 0, out_ptr,
                                              self.be.stream()))

    def block_out_shape(self):
        """(frames per time step, channels, H_out, W_out) of what a stand-alone block returns."""
        s = (ctypes.c_int64 * 4)()
        self._check(self.lib.lu_block_out_shape(self.h, s))
        return tuple(int(v) for v in s)

    def set_graph_mode(self, enable):
        eff = ctypes.c_int32()
        self._check(self.lib.lu_set_graph_mode(self.h, 1 if enable else 0, ctypes.byref(eff)))
        return bool(eff.value)

    def reset_states(self, mask_ptr):
        self._check(self.lib.lu_reset_states(self.h, mask_ptr, self.be.stream()))

    def reset_level_states(self, level, mask_ptr):
        self._check(self.lib.lu_reset_level_states(self.h, int(level), mask_ptr, self.be.stream()))

    def state_shape(self, level, layer):
        s = (ctypes.c_int64 * 4)()
        self._check(self.lib.lu_state_shape(self.h, level, layer, s))
        return tuple(int(v) for v in s)

    def get_state(self, level, layer, which, out_ptr):
        self._check(self.lib.lu_get_state(self.h, level, layer, which, out_ptr, self.be.stream()))

    def set_state(self, level, layer, which, in_ptr):
        self._check(self.lib.lu_set_state(self.h, level, layer, which, in_ptr, self.be.stream()))

    # ---- training -------------------------------------------------------------------------------------
    def loss_backward(self, labels_ptr, class_weights, loss_ptr, grads_ptr):
        cw = (ctypes.c_float * 3)(*[float(v) for v in class_weights])
        self._check(self.lib.lu_loss_backward(self.h, labels_ptr, cw, loss_ptr, grads_ptr, self.be.stream()))

    def set_grad_bucket_callback(self, fn):
        """fn(offset, count) or None; see lu_set_grad_bucket_callback."""
        self._bucket_cb = _lib.GRAD_BUCKET_FN(lambda off, cnt, user: fn(int(off), int(cnt))) if fn is not None else None
        cb = ctypes.cast(self._bucket_cb, ctypes.c_void_p) if self._bucket_cb is not None else None
        self._check(self.lib.lu_set_grad_bucket_callback(self.h, cb, None))

    def set_bn_sync_callback(self, fn, world_size=1):
        """fn(device_pointer, count): sum the `count` float64 values at `device_pointer` over the ranks, in place, on the
        compute stream; None switches synchronised BatchNorm off.  See lu_set_bn_sync_callback."""
        self._bn_cb = _lib.BN_SYNC_FN(lambda ptr, cnt, user: fn(int(ptr), int(cnt))) if fn is not None else None
        cb = ctypes.cast(self._bn_cb, ctypes.c_void_p) if self._bn_cb is not None else None
        self._check(self.lib.lu_set_bn_sync_callback(self.h, cb, None, int(world_size)))

    def workspace_view(self, ptr, count, dtype):
        """A tensor / array aliasing `count` elements of `dtype` at device pointer `ptr` inside the workspace."""
        nbytes = int(count) * np.dtype(dtype).itemsize
        base = self.be.ptr(self._ws_raw)
        off = int(ptr) - base
        if off < 0 or off + nbytes > self.ws_bytes + 1024:
            raise LuError('pointer outside the workspace')
        raw = self._ws_raw[off:off + nbytes]
        if isinstance(raw, np.ndarray):
            return raw.view(dtype)
        td = {np.float64: self.be.torch.float64, np.float32: self.be.torch.float32}[dtype]
        return raw.view(td)

    def adam_step(self, grads_ptr, m_ptr, v_ptr, lr, step, b1=0.9, b2=0.999, eps=1e-7):
        self._check(self.lib.lu_adam_step(self.h, grads_ptr, m_ptr, v_ptr, lr, b1, b2, eps, int(step), self.be.stream()))

    # ---- introspection --------------------------------------------------------------------------------
    def launch_count(self, reset=False):
        n = ctypes.c_int64()
        self._check(self.lib.lu_launch_count(self.h, ctypes.byref(n), 1 if reset else 0))
        return n.value

    def forward_flops(self, T):
        f = ctypes.c_double()
        self._check(self.lib.lu_forward_flops(self.h, int(T), ctypes.byref(f)))
        return f.value

    def lstm_flops(self, T):
        f = ctypes.c_double()
        self._check(self.lib.lu_lstm_flops(self.h, int(T), ctypes.byref(f)))
        return f.value

    def lstm_kernel_time(self, enable):
        """-> (milliseconds, launches) accumulated since the last call; then switches the event recording."""
        ms, n = ctypes.c_float(), ctypes.c_int32()
        self._check(self.lib.lu_lstm_kernel_time(self.h, 1 if enable else 0, ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def kernel_times(self, enable):
        """-> {class: (milliseconds, launches)} of the tensor-core kernel classes (_lib.KERNEL_CLASSES) accumulated
        since the last call; then switches the event recording."""
        n = len(_lib.KERNEL_CLASSES)
        ms, cnt = (ctypes.c_float * n)(), (ctypes.c_int32 * n)()
        self._check(self.lib.lu_kernel_times(self.h, 1 if enable else 0, ms, cnt))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(_lib.KERNEL_CLASSES)}

    def class_flops(self, T):
        """-> {class: algorithmic FLOPs of one training step over T frames per sample at the bound batch}."""
        n = len(_lib.KERNEL_CLASSES)
        f = (ctypes.c_double * n)()
        self._check(self.lib.lu_class_flops(self.h, int(T), f))
        return {k: float(f[i]) for i, k in enumerate(_lib.KERNEL_CLASSES)}

    def debug_buffer(self, name, kind=0):
        """Test hook: an internal NHWC buffer of layer `name` as a host fp32 array (frames, H, W, C)."""
        shape = (ctypes.c_int64 * 4)()
        self._check(self.lib.lu_debug_buffer(self.h, name.encode(), kind, None, shape, self.be.stream()))
        shp = tuple(int(v) for v in shape)
        buf = self.be.zeros(int(np.prod(shp)), np.float32)
        self._check(self.lib.lu_debug_buffer(self.h, name.encode(), kind, self.be.ptr(buf), shape, self.be.stream()))
        return self.be.to_host(buf).reshape(shp)
