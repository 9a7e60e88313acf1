"""TEST-ONLY: numpy buffer backend + the host build of the kernels (tests/_emu/liblu_emu.so, -DLU_HOST_EMU).

Lets the CPU test-suite exercise the library's host logic (plan, activation-staging tables, weight packing,
epilogue / elementwise arithmetic, state handling) through the same C-ABI and the same LuSession driver the product
uses, in a container without a GPU.  The tcgen05 kernel itself does not exist in this build; the scalar mirror
engine executes the identical tables.  The product package never loads this library."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, 'tests', '_emu')
EMU_LIB = os.path.join(EMU_DIR, 'liblu_emu.so')
CSRC = os.path.join(ROOT, 'lstm_unet_b200', 'csrc')


def build_emu(force=False):
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, 'include', 'lstm_unet_b200.h')]
    if not force and os.path.exists(EMU_LIB) and all(os.path.getmtime(EMU_LIB) >= os.path.getmtime(s) for s in srcs):
        return EMU_LIB
    os.makedirs(EMU_DIR, exist_ok=True)
    cmd = ['g++', '-std=c++17', '-O2', '-DLU_HOST_EMU', '-x', 'c++', '-shared', '-fPIC', '-o', EMU_LIB,
           os.path.join(CSRC, 'lu_api.cu')]
    subprocess.run(cmd, check=True, cwd=CSRC)
    return EMU_LIB


class NumpyBackend:
    name = 'numpy-emu'

    def empty(self, n, dtype):
        return np.empty(int(n), dtype=dtype)

    def zeros(self, n, dtype):
        return np.zeros(int(n), dtype=dtype)

    def ptr(self, buf):
        return buf.ctypes.data

    def stream(self):
        return 0

    def to_device(self, arr, out=None):
        a = np.ascontiguousarray(arr)
        if out is None:
            return a.copy()
        out[...] = a.reshape(out.shape)
        return out

    def to_host(self, buf):
        return np.array(buf, copy=True)

    def synchronize(self):
        pass


def emu_session(net_params, **kw):
    from lstm_unet_b200 import _lib
    from lstm_unet_b200.session import LuSession
    lib = _lib.load_library(build_emu())
    assert lib.lu_is_cuda_build() == 0
    kw.setdefault('engine', 'simt')
    cfg = _lib.make_config(net_params, **kw)
    return LuSession(lib, NumpyBackend(), cfg)


def emu_forward(sess, x, training=False):
    """x: numpy (B,T,...) in the API layout; returns (logits, softmax) numpy arrays."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    B, T = x.shape[0], x.shape[1]
    cfg = sess.cfg
    D = 3
    shape = (B, T, D, cfg.height, cfg.width) if cfg.channels_first else (B, T, cfg.height, cfg.width, D)
    # last_depth comes from the parameter layout
    last = [e for e in sess.layout if e['name'].endswith('kernel') and e['name'].startswith('UpLayers')][-1]
    D = last['shape'][3]
    shape = (B, T, D, cfg.height, cfg.width) if cfg.channels_first else (B, T, cfg.height, cfg.width, D)
    logits = np.zeros(shape, dtype=np.float32)
    softmax = np.zeros(shape, dtype=np.float32)
    sess.forward(x.ctypes.data, T, training, logits.ctypes.data, softmax.ctypes.data)
    return logits, softmax


def emu_block_session(cfg):
    """A stand-alone DownBlock2D / UpBlock2D handle (cfg from _lib.make_down_block_config / make_up_block_config)."""
    from lstm_unet_b200 import _lib
    from lstm_unet_b200.session import LuSession
    lib = _lib.load_library(build_emu())
    return LuSession(lib, NumpyBackend(), cfg)


def emu_block_forward(sess, x, skip=None, training=False):
    """-> the 4-D tensor the block returns (`activ` of DownBlock2D.call, the output of UpBlock2D.call), API layout."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    T = x.shape[1] if skip is None else 1
    n, C, H, W = sess.block_out_shape()
    shape = (n * T, C, H, W) if sess.cfg.channels_first else (n * T, H, W, C)
    out = np.zeros(shape, dtype=np.float32)
    sp = None
    if skip is not None:
        skip = np.ascontiguousarray(skip, dtype=np.float32)
        sp = skip.ctypes.data
    sess.block_forward(x.ctypes.data, sp, T, training, out.ctypes.data)
    return out
