"""Device augmentation of training sequences: host driver of lu_augment_sequence / lu_elastic_coords (lu_aug.cuh), the
arithmetic of the reference reader's per-frame chain (``CTCRAMReaderSequence2D._load_and_enqueue`` and its static helpers,
DataHandeling.py:150-395).  Random numbers are drawn on the host by the caller (data.CTCRAMReaderSequence2D draws them in
the reference's order); this class only moves them to the device and launches.  No CPU fallback."""
import ctypes

import numpy as np

from . import _lib
from .session import LuError, TorchCudaBackend


def gaussian_taps(sigma, truncate=4.0):
    """scipy.ndimage.gaussian_filter's kernel: radius int(truncate*sigma + 0.5), exp(-x^2 / 2 sigma^2) normalised."""
    lw = int(truncate * float(sigma) + 0.5)
    x = np.arange(-lw, lw + 1)
    w = np.exp(-0.5 / (float(sigma) * float(sigma)) * x ** 2)
    return lw, (w / w.sum()).astype(np.float64)


def affine_from_points(pts1, pts2):
    """cv2.getAffineTransform (DataHandeling.py:166): the 2x3 float64 matrix mapping three points onto three points."""
    try:
        import cv2
        return cv2.getAffineTransform(np.float32(pts1), np.float32(pts2))
    except ImportError:
        a = np.zeros((6, 6))
        b = np.zeros(6)
        for i in range(3):
            a[2 * i, 0:2], a[2 * i, 2] = pts1[i], 1
            a[2 * i + 1, 3:5], a[2 * i + 1, 5] = pts1[i], 1
            b[2 * i], b[2 * i + 1] = pts2[i]
        return np.linalg.solve(a, b).reshape(2, 3)


def random_affine(shape_size, alpha_affine, random_state):
    """_get_elastic_affine_matrix_ (DataHandeling.py:150-168) with the caller's RandomState."""
    center_square = np.float32(shape_size) // 2
    square_size = min(shape_size) // 3
    pts1 = np.float32([center_square + square_size, [center_square[0] + square_size, center_square[1] - square_size],
                       center_square - square_size])
    pts2 = pts1 + random_state.uniform(-alpha_affine, alpha_affine, size=pts1.shape).astype(np.float32)
    return affine_from_points(pts1, pts2)


class SequenceAugmenter:
    def __init__(self, _lib_override=None, _backend=None):
        self.lib = _lib_override if _lib_override is not None else _lib.load_library()
        self.be = _backend if _backend is not None else TorchCudaBackend()
        self._ws_key, self._ws = None, None

    def _check(self, rc):
        if rc != 0:
            raise LuError(self.lib.lu_last_error().decode())

    def _f32(self, a):
        if isinstance(a, np.ndarray) or not hasattr(a, 'is_cuda'):
            return self.be.to_device(np.ascontiguousarray(a, dtype=np.float32))
        import torch
        return a.to(device=self.be.device, dtype=torch.float32).contiguous()

    def elastic_coords(self, rand2, alpha, sigma):
        """_get_indices4elastic_transform (DataHandeling.py:183-193): rand2 = (2,H,W) float64 uniform [0,1) fields in
        the reference's draw order (x field, y field) -> device (2,H,W) float64 sampling coordinates (y, x)."""
        rand2 = np.ascontiguousarray(rand2, dtype=np.float64)
        _, H, W = rand2.shape
        lw, taps = gaussian_taps(sigma)
        d_rand, d_taps = self.be.to_device(rand2.reshape(-1)), self.be.to_device(taps)
        tmp, out = self.be.empty(2 * H * W, np.float64), self.be.empty(2 * H * W, np.float64)
        self._check(self.lib.lu_elastic_coords(self.be.ptr(d_rand), self.be.ptr(d_taps), lw, H, W, float(alpha),
                                               self.be.ptr(tmp), self.be.ptr(out), self.be.stream()))
        self._keep = (d_rand, d_taps, tmp)
        return out

    def augment(self, img, seg, contrast=None, brightness=None, affine=None, coords=None, flip=(0, 0), rot90=0,
                randomize=True, out_img=None, out_seg=None):
        """img, seg: (T,H,W) crops (numpy or device).  Returns device (T,Ho,Wo) float32 image and segmentation
        ({-1,0,1,2}); Ho x Wo = W x H for odd rot90.  ``out_img`` / ``out_seg``: flat device buffers to write into."""
        T, H, W = tuple(img.shape)
        if rot90 % 2 and H != W:
            raise ValueError('odd rot90 needs square crops (the reference enqueues fixed (H, W) shapes)')
        elastic = affine is not None
        d_img, d_seg = self._f32(img), self._f32(seg)
        d_c = self._f32(contrast) if randomize else None
        d_b = self._f32(brightness) if randomize else None
        key = (T, H, W)
        if self._ws_key != key:
            nb = ctypes.c_size_t()
            self._check(self.lib.lu_aug_workspace_bytes(T, H, W, ctypes.byref(nb)))
            raw = self.be.empty(nb.value + 256, np.uint8)
            self._ws = (raw, (self.be.ptr(raw) + 255) // 256 * 256, nb.value)
            self._ws_key = key
        oi = out_img if out_img is not None else self.be.empty(T * H * W, np.float32)
        os_ = out_seg if out_seg is not None else self.be.empty(T * H * W, np.float32)
        ap = _lib.lu_aug_params()
        ap.frames, ap.H, ap.W = T, H, W
        ap.randomize, ap.elastic = int(bool(randomize)), int(elastic)
        ap.flip0, ap.flip1, ap.rot90 = int(flip[0]), int(flip[1]), int(rot90) % 4
        m = np.asarray(affine, dtype=np.float64).reshape(-1) if elastic else np.array([1, 0, 0, 0, 1, 0], np.float64)
        for i in range(6):
            ap.affine[i] = float(m[i])
        self._check(self.lib.lu_augment_sequence(
            self.be.ptr(d_img), self.be.ptr(d_seg), self.be.ptr(d_c) if randomize else None,
            self.be.ptr(d_b) if randomize else None, self.be.ptr(coords) if elastic else None, ctypes.byref(ap),
            self.be.ptr(oi), self.be.ptr(os_), self._ws[1], self._ws[2], self.be.stream()))
        self._keep2 = (d_img, d_seg, d_c, d_b, coords)
        shape = (T, W, H) if rot90 % 2 else (T, H, W)
        return oi.reshape(shape), os_.reshape(shape)
