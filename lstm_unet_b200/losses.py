"""Mirror of the reference's ``losses.WeightedCELoss`` (losses.py:8-27) for the CUDA model.

The reference computes the loss from the logits tensor the model returned; here the logits of the last model call are
still resident in the library's workspace (fp32, exactly the values returned), so the loss kernel reads them there.
``seg_measure`` (losses.py:29-88: a tf.py_function with SciPy labelling and Python loops on the host, called inside
every train / validation step, train2D.py:97,111) runs on the device (lu_seg_measure, lu_post.cuh) from the labels and
logits that are already there; the same launch sequence also yields the sparse categorical accuracy of
train2D.py:98-101."""
import ctypes

import numpy as np


class WeightedCELoss(object):
    def __init__(self, channel_axis, class_weights):
        self.channel_axis = channel_axis
        self.class_weights = class_weights

    def __call__(self, gt_sequence, output_sequence):
        model = getattr(output_sequence, '_lu_model', None)
        if model is None:
            raise ValueError('WeightedCELoss expects the logits returned by the last ULSTMnet2D call')
        return model.loss(gt_sequence, self.class_weights)


class _StepMetrics:
    """Device SEG measure + accuracy of (labels, logits); one workspace per frame shape."""

    def __init__(self, channel_axis, _lib_override=None, _backend=None):
        from . import _lib
        from .session import TorchCudaBackend
        if channel_axis not in (2, 4):
            raise ValueError('channel_axis of the 5-D (B,T,...) tensors must be 2 (NCHW) or 4 (NHWC)')
        self.channels_first = channel_axis == 2
        self.lib = _lib_override if _lib_override is not None else _lib.load_library()
        self.be = _backend if _backend is not None else TorchCudaBackend()
        self._ws_key, self._ws = None, None

    def _dev(self, a):
        if isinstance(a, np.ndarray):
            return self.be.to_device(np.ascontiguousarray(a, dtype=np.float32))
        if hasattr(a, 'is_cuda'):
            import torch
            a = a.as_subclass(torch.Tensor)
            if not a.is_cuda or a.dtype != torch.float32 or not a.is_contiguous():
                a = a.to(device=self.be.device, dtype=torch.float32).contiguous()
        return a

    def compute(self, gt_sequence, output_sequence):
        """-> (SEG, accuracy) as Python floats (one 32-byte device->host read)."""
        from .session import LuError
        lab, lg = self._dev(gt_sequence), self._dev(output_sequence)
        shp = tuple(lg.shape)
        if len(shp) != 5:
            raise ValueError('expected 5-D (B,T,...) logits')
        if self.channels_first:
            (B, T, C, H, W) = shp
        else:
            (B, T, H, W, C) = shp
        if C != 3 or int(np.prod(tuple(lab.shape))) != B * T * H * W:
            raise ValueError('labels %s do not match logits %s' % (tuple(lab.shape), shp))
        n = B * T
        key = (n, H, W)
        if self._ws_key != key:
            nb = ctypes.c_size_t()
            if self.lib.lu_seg_workspace_bytes(n, H, W, ctypes.byref(nb)):
                raise LuError(self.lib.lu_last_error().decode())
            raw = self.be.empty(nb.value + 256, np.uint8)
            self._ws = (raw, (self.be.ptr(raw) + 255) // 256 * 256, nb.value)
            self._ws_key = key
        res = self.be.zeros(4, np.float64)
        if self.lib.lu_seg_measure(self.be.ptr(lab), self.be.ptr(lg), n, H, W, 1 if self.channels_first else 0,
                                   self.be.ptr(res), self._ws[1], self._ws[2], self.be.stream()):
            raise LuError(self.lib.lu_last_error().decode())
        r = np.asarray(self.be.to_host(res), dtype=np.float64)
        seg = float(r[0] / r[1]) if r[1] > 0 else float('nan')
        return seg, float(r[2] / r[3])


def seg_measure(channel_axis, three_d=False, foreground_class_index=1, **kw):
    """losses.seg_measure (losses.py:29-88): returns ``calc_seg(gt_sequence, output_sequence) -> SEG`` for 5-D
    (B,T,...) tensors.  ``calc_seg.last_accuracy`` holds the sparse categorical accuracy of the same call."""
    if three_d:
        raise ValueError('only the 2-D path (ULSTMnet2D) is built')
    if foreground_class_index != 1:
        raise ValueError('foreground_class_index is 1 in the reference\'s use (train2D.py:51)')
    m = _StepMetrics(channel_axis, **kw)

    def calc_seg(gt_sequence, output_sequence):
        seg, acc = m.compute(gt_sequence, output_sequence)
        calc_seg.last_accuracy = acc
        return seg
    calc_seg.last_accuracy = None
    return calc_seg
