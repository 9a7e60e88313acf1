"""ctypes binding of liblstm_unet_b200.so (C-ABI declared in include/lstm_unet_b200.h).

There is no CPU fallback: `load_library()` raises if the CUDA library is missing or is not the CUDA build.
"""
import ctypes
import os

LU_MAX_LEVELS = 4
LU_MAX_PER_LEVEL = 4
PRECISIONS = {'bf16': 0, 'bf16x3': 1, 'fp16': 2}
ENGINES = {'tcgen05': 0, 'simt': 1}
GATES = {'hard_sigmoid': 0, 'sigmoid': 1}
KERNEL_CLASSES = ('lstm_fwd', 'conv_fwd', 'dgrad', 'wgrad')     # LU_KC_* of include/lstm_unet_b200.h
A_MODES = {'halo': 0, 'direct': 1}
BLOCK_NET, BLOCK_DOWN, BLOCK_UP = 0, 1, 2                      # LU_BLOCK_*

_I32 = ctypes.c_int32
_A1 = _I32 * LU_MAX_LEVELS
_A2 = (_I32 * LU_MAX_PER_LEVEL) * LU_MAX_LEVELS


class lu_config(ctypes.Structure):
    _fields_ = [
        ('n_levels', _I32),
        ('n_lstm', _A1), ('lstm_k', _A2), ('lstm_f', _A2),
        ('n_down', _A1), ('down_k', _A2), ('down_f', _A2),
        ('n_up', _A1), ('up_k', _A2), ('up_f', _A2),
        ('in_channels', _I32), ('channels_first', _I32), ('pad_image', _I32),
        ('batch', _I32), ('max_t', _I32), ('height', _I32), ('width', _I32),
        ('precision', _I32), ('engine', _I32), ('gate', _I32), ('a_mode', _I32), ('train', _I32),
        ('lrelu_alpha', ctypes.c_float),
        ('block_kind', _I32), ('block_stride', _I32), ('skip_channels', _I32), ('return_logits', _I32),
    ]


class lu_post_params(ctypes.Structure):
    _fields_ = [('edge_thresh', ctypes.c_float), ('edge_d2_limit', _I32), ('min_cell_size', _I32),
                ('max_cell_size', _I32), ('fov', _I32), ('channels_first', _I32)]


class lu_aug_params(ctypes.Structure):
    _fields_ = [('frames', _I32), ('H', _I32), ('W', _I32), ('randomize', _I32), ('elastic', _I32), ('flip0', _I32),
                ('flip1', _I32), ('rot90', _I32), ('affine', ctypes.c_double * 6)]


GRAD_BUCKET_FN = ctypes.CFUNCTYPE(None, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p)
BN_SYNC_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)

LIB_NAME = 'liblstm_unet_b200.so'


def default_library_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


def bind(lib):
    """Declare argument / return types of every entry point of include/lstm_unet_b200.h."""
    vp, i32, i64, f32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float
    P = ctypes.POINTER
    lib.lu_last_error.restype = ctypes.c_char_p
    lib.lu_last_error.argtypes = []
    sigs = {
        'lu_version': [], 'lu_is_cuda_build': [],
        'lu_create': [P(lu_config), P(vp)],
        'lu_destroy': [vp],
        'lu_workspace_bytes': [vp, P(ctypes.c_size_t)],
        'lu_bind_workspace': [vp, vp, ctypes.c_size_t, vp],
        'lu_param_count': [vp, P(i32), P(i64), P(i64)],
        'lu_param_info': [vp, i32, ctypes.c_char_p, i32, P(i64), P(i32), P(i64), P(i32)],
        'lu_bind_params': [vp, vp],
        'lu_params_changed': [vp, vp],
        'lu_forward': [vp, vp, i32, i32, vp, vp, vp],
        'lu_block_forward': [vp, vp, vp, i32, i32, vp, vp],
        'lu_block_out_shape': [vp, P(i64)],
        'lu_set_graph_mode': [vp, i32, P(i32)],
        'lu_reset_states': [vp, vp, vp],
        'lu_reset_level_states': [vp, i32, vp, vp],
        'lu_state_shape': [vp, i32, i32, P(i64)],
        'lu_get_state': [vp, i32, i32, i32, vp, vp],
        'lu_set_state': [vp, i32, i32, i32, vp, vp],
        'lu_loss_backward': [vp, vp, P(f32), vp, vp, vp],
        'lu_set_grad_bucket_callback': [vp, vp, vp],
        'lu_set_bn_sync_callback': [vp, vp, vp, i32],
        'lu_adam_step': [vp, vp, vp, vp, f32, f32, f32, f32, i64, vp],
        'lu_debug_buffer': [vp, ctypes.c_char_p, i32, vp, P(i64), vp],
        'lu_launch_count': [vp, P(i64), i32],
        'lu_forward_flops': [vp, i32, P(ctypes.c_double)],
        'lu_lstm_flops': [vp, i32, P(ctypes.c_double)],
        'lu_lstm_kernel_time': [vp, i32, P(f32), P(i32)],
        'lu_kernel_times': [vp, i32, P(f32), P(i32)],
        'lu_class_flops': [vp, i32, P(ctypes.c_double)],
        'lu_post_workspace_bytes': [i32, i32, i32, P(ctypes.c_size_t)],
        'lu_postprocess': [vp, i32, i32, i32, P(lu_post_params), vp, vp, vp, ctypes.c_size_t, vp],
        'lu_post_launch_count': [P(i64), i32],
        'lu_seg_workspace_bytes': [i32, i32, i32, P(ctypes.c_size_t)],
        'lu_seg_measure': [vp, vp, i32, i32, i32, i32, vp, vp, ctypes.c_size_t, vp],
        'lu_aug_workspace_bytes': [i32, i32, i32, P(ctypes.c_size_t)],
        'lu_augment_sequence': [vp, vp, vp, vp, vp, P(lu_aug_params), vp, vp, vp, ctypes.c_size_t, vp],
        'lu_elastic_coords': [vp, vp, i32, i32, i32, ctypes.c_double, vp, vp, vp],
    }
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ctypes.c_int
    return lib


EXPORTED_SYMBOLS = ['lu_last_error', 'lu_version', 'lu_is_cuda_build', 'lu_create', 'lu_destroy', 'lu_workspace_bytes',
                    'lu_bind_workspace', 'lu_param_count', 'lu_param_info', 'lu_bind_params', 'lu_params_changed',
                    'lu_forward', 'lu_block_forward', 'lu_block_out_shape', 'lu_set_graph_mode', 'lu_reset_states', 'lu_reset_level_states', 'lu_state_shape', 'lu_get_state', 'lu_set_state',
                    'lu_loss_backward', 'lu_set_grad_bucket_callback', 'lu_set_bn_sync_callback', 'lu_adam_step', 'lu_debug_buffer', 'lu_launch_count', 'lu_forward_flops', 'lu_lstm_flops',
                    'lu_lstm_kernel_time', 'lu_kernel_times', 'lu_class_flops', 'lu_post_workspace_bytes', 'lu_postprocess', 'lu_post_launch_count',
                    'lu_seg_workspace_bytes', 'lu_seg_measure', 'lu_aug_workspace_bytes', 'lu_augment_sequence',
                    'lu_elastic_coords']

_LIB = None


def load_library(path=None):
    """Load the CUDA library.  Raises RuntimeError (never falls back) if it is missing or not the CUDA build."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or default_library_path()
    if not os.path.exists(p):
        raise RuntimeError('%s not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                           '(nvcc, sm_100a). There is no CPU fallback.' % p)
    lib = bind(ctypes.CDLL(p))
    if path is None:
        if lib.lu_is_cuda_build() != 1:
            raise RuntimeError('%s is not the CUDA product build' % p)
        _LIB = lib
    return lib


def make_config(net_params, data_format='NCHW', pad_image=True, batch=1, max_t=1, height=0, width=0,
                precision='bf16', engine='tcgen05', gate='hard_sigmoid', a_mode='halo', train=False, in_channels=1,
                lrelu_alpha=0.3):
    """Flatten the reference's `net_kernel_params` dict (Params.py:49-69) + call shapes into the C struct.
    Raises ValueError on the level-count mismatches ULSTMnet2D.__init__ rejects (Networks.py:188-193)."""
    down, lstm, up = net_params['down_conv_kernels'], net_params['lstm_kernels'], net_params['up_conv_kernels']
    if not len(down) == len(lstm):
        raise ValueError('Number of layers in down path ({}) do not match number of LSTM layers ({})'.format(
            len(down), len(lstm)))
    if not len(down) == len(up):
        raise ValueError('Number of layers in down path ({}) do not match number of layers in up path ({})'.format(
            len(down), len(up)))
    if len(down) > LU_MAX_LEVELS:
        raise ValueError('at most %d levels are supported' % LU_MAX_LEVELS)
    c = lu_config()
    c.n_levels = len(down)
    for name_n, name_k, name_f, lists in (('n_lstm', 'lstm_k', 'lstm_f', lstm), ('n_down', 'down_k', 'down_f', down),
                                          ('n_up', 'up_k', 'up_f', up)):
        for li, layer in enumerate(lists):
            if len(layer) > LU_MAX_PER_LEVEL:
                raise ValueError('at most %d layers per level are supported' % LU_MAX_PER_LEVEL)
            getattr(c, name_n)[li] = len(layer)
            for j, (k, f) in enumerate(layer):
                getattr(c, name_k)[li][j] = int(k)
                getattr(c, name_f)[li][j] = int(f)
    c.in_channels = in_channels
    c.channels_first = 1 if data_format[1] == 'C' else 0      # Networks.py:181-182
    c.pad_image = 1 if pad_image else 0
    c.batch, c.max_t, c.height, c.width = int(batch), int(max_t), int(height), int(width)
    c.precision = PRECISIONS[precision]
    c.engine = ENGINES[engine]
    c.gate = GATES[gate]
    c.a_mode = A_MODES[a_mode]
    c.train = 1 if train else 0
    c.lrelu_alpha = float(lrelu_alpha)      # Keras-2 LeakyReLU() default (Networks.py:58,139)
    return c


def _common(c, data_format, batch, max_t, height, width, precision, engine, gate, a_mode, in_channels, lrelu_alpha):
    c.in_channels = int(in_channels)
    c.channels_first = 1 if data_format[1] == 'C' else 0
    c.pad_image = 0
    c.batch, c.max_t, c.height, c.width = int(batch), int(max_t), int(height), int(width)
    c.precision, c.engine, c.gate, c.a_mode = PRECISIONS[precision], ENGINES[engine], GATES[gate], A_MODES[a_mode]
    c.train = 0
    c.lrelu_alpha = float(lrelu_alpha)
    return c


def _fill(c, prefix, layer):
    if len(layer) > LU_MAX_PER_LEVEL:
        raise ValueError('at most %d layers per block are supported' % LU_MAX_PER_LEVEL)
    getattr(c, 'n_' + prefix)[0] = len(layer)
    for j, (k, f) in enumerate(layer):
        getattr(c, prefix + '_k')[0][j] = int(k)
        getattr(c, prefix + '_f')[0][j] = int(f)


def make_down_block_config(conv_kernels, lstm_kernels, stride=2, data_format='NCHW', batch=1, max_t=1, height=0, width=0,
                           in_channels=1, precision='bf16', engine='tcgen05', gate='hard_sigmoid', a_mode='halo',
                           lrelu_alpha=0.3):
    """DownBlock2D(conv_kernels, lstm_kernels, stride, data_format) on its own (Networks.py:37-75)."""
    c = lu_config()
    c.n_levels = 1
    _fill(c, 'lstm', lstm_kernels)
    _fill(c, 'down', conv_kernels)
    c.block_kind, c.block_stride = BLOCK_DOWN, int(stride)
    return _common(c, data_format, batch, max_t, height, width, precision, engine, gate, a_mode, in_channels, lrelu_alpha)


def make_up_block_config(kernels, up_factor=2, data_format='NCHW', return_logits=False, frames=1, height=0, width=0,
                         in_channels=1, skip_channels=1, precision='bf16', engine='tcgen05', a_mode='halo',
                         lrelu_alpha=0.3):
    """UpBlock2D(kernels, up_factor, data_format, return_logits) on its own (Networks.py:124-153); height / width /
    in_channels describe the low-resolution input."""
    c = lu_config()
    c.n_levels = 1
    _fill(c, 'up', kernels)
    c.block_kind, c.block_stride = BLOCK_UP, int(up_factor)
    c.skip_channels, c.return_logits = int(skip_channels), 1 if return_logits else 0
    return _common(c, data_format, frames, 1, height, width, precision, engine, 'hard_sigmoid', a_mode, in_channels,
                   lrelu_alpha)
