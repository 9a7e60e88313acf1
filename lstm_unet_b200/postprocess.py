"""Instance labelling of the model's soft-max on the device -- the step that follows the model call in the
reference's ``Inference2D.inference`` (Inference2D.py:64-123; parameters: ``CTCInferenceParams.edge_dist``,
``min_cell_size``, ``max_cell_size``, ``FOV``, Params.py:164-167).

    reference (numpy / SciPy / OpenCV, per frame, on the host)     here
    ---------------------------------------------------------------------------------------------------------------
    image_softmax.numpy() -> threshold, fill holes, connected      PostProcessor(params)(softmax) -> uint16 labels,
    components, EDT edge assignment, per-cell hole loop, FOV       computed by lu_postprocess (lu_post.cuh) from the
    and size filters -> labels_out (uint16)                        soft-max that is already in HBM; only the label
                                                                   image (2 bytes / pixel) crosses PCIe

Results are bit-identical to the reference's (tests/test_gpu_postprocess.py, tests/golden/postprocess.npz).  There is
no CPU fallback: without the CUDA library / a CUDA device the constructor raises.
"""
import ctypes
import math

import numpy as np

from . import _lib
from .session import LuError, TorchCudaBackend


def edge_d2_limit(edge_dist):
    """Exclusive bound on the squared pixel distance equivalent to the reference's float64 test
    ``distance_transform_edt(...) < params.edge_dist`` (Inference2D.py:78): smallest d2 with sqrt(d2) >= edge_dist."""
    e = float(edge_dist)
    if not e > 0:
        return 0
    d2 = max(0, int(math.floor(e * e)) - 2)
    while math.sqrt(float(d2)) < e:
        d2 += 1
    return d2


class PostProcessor:
    """``PostProcessor(params)`` with a ``CTCInferenceParams``-like object, or keyword overrides.  Calling it with the
    soft-max of N frames -- (3,H,W), (N,3,H,W) or the model's (B,T,3,H,W) output, device tensor or numpy -- returns the
    uint16 label images with the same leading dimensions (device tensor; ``.numpy()`` like the model outputs)."""

    def __init__(self, params=None, edge_dist=None, min_cell_size=None, max_cell_size=None, FOV=None, edge_thresh=0.2,
                 data_format=None, _lib_override=None, _backend=None):
        def pick(v, name, default):
            if v is not None:
                return v
            return getattr(params, name, default) if params is not None else default
        self.edge_dist = pick(edge_dist, 'edge_dist', 2)
        self.min_cell_size = int(pick(min_cell_size, 'min_cell_size', 10))
        self.max_cell_size = int(pick(max_cell_size, 'max_cell_size', 100))
        self.FOV = int(pick(FOV, 'FOV', 0))
        self.edge_thresh = float(edge_thresh)
        fmt = pick(data_format, 'data_format', 'NCHW')
        self.channels_first = fmt[1] == 'C'
        self.lib = _lib_override if _lib_override is not None else _lib.load_library()
        self.be = _backend if _backend is not None else TorchCudaBackend()
        self._ws = None
        self._ws_key = None
        self.last_info = None

    def _check(self, rc):
        if rc != 0:
            raise LuError(self.lib.lu_last_error().decode())

    def _workspace(self, n, H, W):
        key = (n, H, W)
        if self._ws_key != key:
            nb = ctypes.c_size_t()
            self._check(self.lib.lu_post_workspace_bytes(n, H, W, ctypes.byref(nb)))
            raw = self.be.empty(nb.value + 256, np.uint8)
            base = self.be.ptr(raw)
            self._ws = (raw, (base + 255) // 256 * 256, nb.value)
            self._ws_key = key
        return self._ws

    def __call__(self, softmax):
        be = self.be
        if isinstance(softmax, np.ndarray):
            sm = be.to_device(np.ascontiguousarray(softmax, dtype=np.float32))
        else:
            sm = softmax
            if hasattr(sm, 'is_cuda'):
                import torch
                sm = sm.as_subclass(torch.Tensor)
                if not sm.is_cuda or sm.dtype != torch.float32 or not sm.is_contiguous():
                    sm = sm.to(device=be.device, dtype=torch.float32).contiguous()
        shape = tuple(sm.shape)
        if len(shape) < 3:
            raise ValueError('soft-max must be (..., 3, H, W) or (..., H, W, 3)')
        if self.channels_first:
            lead, (C, H, W) = shape[:-3], shape[-3:]
        else:
            lead, (H, W, C) = shape[:-3], shape[-3:]
        if C != 3:
            raise ValueError('soft-max must have 3 classes (background, cell, edge), got %d' % C)
        n = int(np.prod(lead)) if lead else 1
        _, ws_ptr, ws_bytes = self._workspace(n, H, W)
        out = be.empty(n * H * W * 2, np.uint8)
        info = be.empty(n * 16, np.uint8)
        pp = _lib.lu_post_params(self.edge_thresh, edge_d2_limit(self.edge_dist), self.min_cell_size,
                                 self.max_cell_size, self.FOV, 1 if self.channels_first else 0)
        self._check(self.lib.lu_postprocess(be.ptr(sm), n, H, W, ctypes.byref(pp), be.ptr(out), be.ptr(info), ws_ptr,
                                            ws_bytes, be.stream()))
        self._keepalive = sm
        self.last_info = info
        return self._as_labels(out, lead + (H, W))

    def _as_labels(self, out, shape):
        if isinstance(out, np.ndarray):
            return out.view(np.uint16).reshape(shape)
        from .Networks import _wrap
        return _wrap(out.view(self.be.torch.uint16).reshape(shape))

    def info(self):
        """int32 (frames, 4) of the last call: components incl. background (cv2's count), labels kept, sequential pass
        needed, 0."""
        if self.last_info is None:
            return None
        a = self.last_info if isinstance(self.last_info, np.ndarray) else self.be.to_host(self.last_info)
        return np.asarray(a).view(np.int32).reshape(-1, 4)
