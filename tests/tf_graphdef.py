"""ULSTMnet2D as a TensorFlow GraphDef, written without TensorFlow, for execution by a THIRD-PARTY TensorFlow runtime.

TensorFlow cannot be installed here, but OpenCV ships its own implementation of the TensorFlow operators
(`cv2.dnn.readNetFromTensorflow`, built to reproduce TensorFlow's outputs on frozen graphs).  This module serialises the
forward of the reference network as the graph of TF ops Keras-2 lowers it to -- `Conv2D` (padding "SAME") + `BiasAdd`,
`FusedBatchNormV3` (is_training = false, epsilon = 1e-3), `LeakyRelu` (alpha = 0.3), `ResizeBilinear`
(half_pixel_centers), `ConcatV2`, `MirrorPad` (REFLECT), `Softmax`, and the ConvLSTM2D cell as Keras-2's
`ConvLSTM2DCell.call` writes it (kernel split in four, eight convolutions, hard_sigmoid = clip(0.2 x + 0.5, 0, 1),
gate order i, f, c, o) -- so that an implementation of those operators that is neither the oracle's (torch) nor this
repo's (CUDA) can be put beside both.  What it pins: the TensorFlow operator semantics of SURVEY App. A (asymmetric SAME
padding under stride 2, where epsilon enters the fused batch norm, half-pixel bilinear sampling, REFLECT padding).  What
it cannot pin: that Keras-2 lowers the reference's layers to exactly this graph (that part is restated from the Keras-2
sources: Networks.py:48-58,135-145,206,232 name the layers, this file names the ops).

The protobuf wire format is written by hand (field numbers of tensorflow/core/framework/{graph,node_def,attr_value,
tensor,tensor_shape,types}.proto); the only consumer is the test suite.
"""
import struct

import numpy as np

DT_FLOAT, DT_INT32 = 1, 3


# ---- protobuf wire format ---------------------------------------------------------------------------------------------
def _varint(n):
    if n < 0:
        n += 1 << 64
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _tag(field, wire):
    return _varint((field << 3) | wire)


def f_varint(field, v):
    return _tag(field, 0) + _varint(int(v))


def f_bytes(field, b):
    if isinstance(b, str):
        b = b.encode()
    return _tag(field, 2) + _varint(len(b)) + b


def f_float(field, v):
    return _tag(field, 5) + struct.pack('<f', v)


# ---- TensorFlow messages ----------------------------------------------------------------------------------------------
def shape_proto(dims):                       # TensorShapeProto: repeated Dim dim = 2 { int64 size = 1 }
    return b''.join(f_bytes(2, f_varint(1, d)) for d in dims)


def tensor_proto(arr):                       # TensorProto: dtype = 1, tensor_shape = 2, tensor_content = 4
    arr = np.ascontiguousarray(arr)
    assert arr.dtype in (np.float32, np.int32)
    dt = DT_FLOAT if arr.dtype == np.float32 else DT_INT32
    return f_varint(1, dt) + f_bytes(2, shape_proto(arr.shape)) + f_bytes(4, arr.tobytes())


# AttrValue: list = 1, s = 2, i = 3, f = 4, b = 5, type = 6, shape = 7, tensor = 8
def attr_s(s): return f_bytes(2, s)
def attr_i(i): return f_varint(3, i)
def attr_f(v): return f_float(4, v)
def attr_b(b): return f_varint(5, 1 if b else 0)
def attr_type(t): return f_varint(6, t)
def attr_shape(dims): return f_bytes(7, shape_proto(dims))
def attr_tensor(arr): return f_bytes(8, tensor_proto(arr))
def attr_list_i(vals): return f_bytes(1, f_bytes(3, b''.join(_varint(v) for v in vals)))   # ListValue.i = 3, packed


def node(name, op, inputs=(), **attrs):      # GraphDef.node = 1; NodeDef: name = 1, op = 2, input = 3, attr = 5 (map)
    body = f_bytes(1, name) + f_bytes(2, op) + b''.join(f_bytes(3, i) for i in inputs)
    for k, v in attrs.items():
        body += f_bytes(5, f_bytes(1, k) + f_bytes(2, v))
    return f_bytes(1, body)


_F = attr_type(DT_FLOAT)


def const(name, arr):
    arr = np.asarray(arr)
    return node(name, 'Const', dtype=attr_type(DT_FLOAT if arr.dtype == np.float32 else DT_INT32), value=attr_tensor(arr))


def placeholder(name, dims):
    return node(name, 'Placeholder', dtype=_F, shape=attr_shape(dims))


# ---- the operators Keras-2 lowers the reference's layers to --------------------------------------------------------------
def conv2d(name, x, w_hwio, stride=1, bias=None):
    """Conv2D(padding='same') [+ BiasAdd]: keras Conv2D.call -> nn.conv2d + nn.bias_add."""
    w = np.ascontiguousarray(w_hwio, dtype=np.float32)
    cname = name + '/Conv2D' if bias is not None else name
    g = const(name + '/kernel', w)
    g += node(cname, 'Conv2D', [x, name + '/kernel'], T=_F, strides=attr_list_i([1, stride, stride, 1]),
              padding=attr_s('SAME'), data_format=attr_s('NHWC'), dilations=attr_list_i([1, 1, 1, 1]))
    if bias is not None:
        g += const(name + '/bias', np.ascontiguousarray(bias, dtype=np.float32))
        g += node(name, 'BiasAdd', [cname, name + '/bias'], T=_F, data_format=attr_s('NHWC'))
    return g


def select_channels(name, x, c_total, c0, n):
    """x[..., c0:c0+n] as a 1x1 convolution with a 0/1 kernel (exact in fp32): keeps every tensor of the graph downstream
    of ONE placeholder and of a Conv2D, which is what OpenCV's importer needs to track the NHWC layout."""
    w = np.zeros((1, 1, c_total, n), np.float32)
    for j in range(n):
        w[0, 0, c0 + j, j] = 1.0
    return conv2d(name, x, w)


def zeros_like_channels(name, x, c_total, n):
    return conv2d(name, x, np.zeros((1, 1, c_total, n), np.float32))


def batchnorm(name, x, gamma, beta, mean, var, eps=1e-3):
    g = b''
    for suffix, a in (('gamma', gamma), ('beta', beta), ('moving_mean', mean), ('moving_variance', var)):
        g += const(name + '/' + suffix, np.ascontiguousarray(a, dtype=np.float32))
    g += node(name, 'FusedBatchNormV3', [x] + [name + '/' + s for s in ('gamma', 'beta', 'moving_mean', 'moving_variance')],
              T=_F, U=_F, epsilon=attr_f(eps), data_format=attr_s('NHWC'), is_training=attr_b(False))
    return g


def leaky_relu(name, x, alpha=0.3):
    return node(name, 'LeakyRelu', [x], T=_F, alpha=attr_f(alpha))


def resize_bilinear(name, x, out_h, out_w):
    """keras.backend.resize_images(..., interpolation='bilinear') -> tf.image.resize (v2): half-pixel centres."""
    return const(name + '/size', np.array([out_h, out_w], np.int32)) + \
        node(name, 'ResizeBilinear', [x, name + '/size'], T=_F, align_corners=attr_b(False), half_pixel_centers=attr_b(True))


def const_int_scalar(name, v):
    """Scalar int32 Const the way TensorFlow itself serialises it: TensorProto.int_val = 7 (OpenCV's importer reads the
    concat axis from there and does not look at tensor_content)."""
    t = f_varint(1, DT_INT32) + f_bytes(2, b'') + f_varint(7, v)
    return node(name, 'Const', dtype=attr_type(DT_INT32), value=f_bytes(8, t))


def concat(name, xs):
    return const_int_scalar(name + '/axis', 3) + \
        node(name, 'ConcatV2', list(xs) + [name + '/axis'], T=_F, N=attr_i(len(xs)), Tidx=attr_type(DT_INT32))


def mirror_pad(name, x, pt, pb, pl, pr):
    return const(name + '/paddings', np.array([[0, 0], [pt, pb], [pl, pr], [0, 0]], np.int32)) + \
        node(name, 'MirrorPad', [x, name + '/paddings'], T=_F, Tpaddings=attr_type(DT_INT32), mode=attr_s('REFLECT'))


def softmax(name, x):
    return node(name, 'Softmax', [x], T=_F)


def hard_sigmoid(name, x):
    """keras.backend.hard_sigmoid (TF backend, Keras 2): x * 0.2 + 0.5, then clip_by_value(0, 1)."""
    g = const(name + '/point_two', np.array(0.2, np.float32)) + const(name + '/point_five', np.array(0.5, np.float32))
    g += const(name + '/zero', np.array(0.0, np.float32)) + const(name + '/one', np.array(1.0, np.float32))
    g += node(name + '/mul', 'Mul', [x, name + '/point_two'], T=_F)
    g += node(name + '/add', 'AddV2', [name + '/mul', name + '/point_five'], T=_F)
    g += node(name + '/min', 'Minimum', [name + '/add', name + '/one'], T=_F)
    g += node(name, 'Maximum', [name + '/min', name + '/zero'], T=_F)
    return g


def convlstm_cell(name, x, h, c, kernel, recurrent_kernel, bias):
    """One step of keras ConvLSTM2DCell.call (Keras-2 defaults: padding same, tanh / hard_sigmoid, use_bias): the kernels are
    split in four along the output axis (i, f, c, o), eight convolutions, the recurrent ones without bias.
    Returns (graph bytes, h_name, c_name)."""
    F_ = recurrent_kernel.shape[2]
    g = b''
    for gi, gn in enumerate('ifco'):
        sl = slice(gi * F_, (gi + 1) * F_)
        g += conv2d('%s/x_%s' % (name, gn), x, kernel[..., sl], 1, bias[sl])
        g += conv2d('%s/h_%s' % (name, gn), h, recurrent_kernel[..., sl], 1, None)
        g += node('%s/z_%s' % (name, gn), 'AddV2', ['%s/x_%s' % (name, gn), '%s/h_%s' % (name, gn)], T=_F)
    for gn in 'ifo':
        g += hard_sigmoid('%s/%s' % (name, gn), '%s/z_%s' % (name, gn))
    g += node(name + '/g', 'Tanh', [name + '/z_c'], T=_F)
    g += node(name + '/f_c', 'Mul', [name + '/f', c], T=_F) + node(name + '/i_g', 'Mul', [name + '/i', name + '/g'], T=_F)
    g += node(name + '/c', 'AddV2', [name + '/f_c', name + '/i_g'], T=_F)
    g += node(name + '/tanh_c', 'Tanh', [name + '/c'], T=_F)
    g += node(name + '/h', 'Mul', [name + '/o', name + '/tanh_c'], T=_F)
    return g, name + '/h', name + '/c'


def build_ulstm_graph(net_params, params, B, T, C, H, W, pad_image):
    """ULSTMnet2D.call (Networks.py:208-254) for a (B, T, H, W, C) sequence, inference mode, unrolled over T from zero
    ConvLSTM states (the first call of the stateful layers; k stateful calls of T frames == one call of k*T frames here,
    batch-norm being frozen).

    One placeholder 'inp' of shape (B, H, W, T*C): the frames stacked along the channel axis.  Returns (graph bytes,
    output names): the un-cropped logits / soft-max of every frame and the final h / c of every ConvLSTM layer; the crop
    of Networks.py:250 is index arithmetic and is applied by the caller with the returned pads."""
    n_levels = len(net_params['down_conv_kernels'])
    total_stride = 2 ** (n_levels - 1)
    min_pad = total_stride if pad_image else 0
    pad_y = (min_pad, min_pad + (total_stride - H % total_stride) % total_stride)
    pad_x = (min_pad, min_pad + (total_stride - W % total_stride) % total_stride)
    Hp, Wp = H + sum(pad_y), W + sum(pad_x)
    p = {k: np.asarray(v, dtype=np.float32) for k, v in params.items()}

    c_total = T * C
    g = placeholder('inp', [B, H, W, c_total])
    src = 'inp'
    if max(pad_y) or max(pad_x):
        g += mirror_pad('pad', 'inp', pad_y[0], pad_y[1], pad_x[0], pad_x[1])      # tf.pad(..., 'REFLECT') (:232)
        src = 'pad'
    frames = []
    for t in range(T):
        g += select_channels('frame_%d' % t, src, c_total, t * C, C)
        frames.append('frame_%d' % t)

    skips_per_t = [[] for _ in range(T)]
    cur, cin = frames, C
    final_states = []
    hh, ww = Hp, Wp
    for li in range(n_levels):
        for t in range(T):
            skips_per_t[t].append((cur[t], cin))                                    # skip = the block's INPUT (:239)
        # ConvLSTM2D layers, unrolled over time (stateful: h, c thread through t)
        for j, (k, f) in enumerate(net_params['lstm_kernels'][li]):
            pre = 'DownLayers/%d/ConvLSTM/%d/' % (li, j)
            g += zeros_like_channels(pre + 'h_init', cur[0], cin, f) + zeros_like_channels(pre + 'c_init', cur[0], cin, f)
            h, c = pre + 'h_init', pre + 'c_init'
            outs = []
            for t in range(T):
                gg, h, c = convlstm_cell('%st%d' % (pre, t), cur[t], h, c, p[pre + 'kernel'], p[pre + 'recurrent_kernel'],
                                         p[pre + 'bias'])
                g += gg
                outs.append(h)
            final_states.append((h, c))
            cur, cin = outs, f
        stride = 2 if li < n_levels - 1 else 1
        for j, (k, f) in enumerate(net_params['down_conv_kernels'][li]):
            pre = 'DownLayers/%d/' % li
            nxt = []
            for t in range(T):
                nm = '%st%d/' % (pre, t)
                g += conv2d(nm + 'Conv/%d' % j, cur[t], p[pre + 'Conv/%d/kernel' % j], stride if j == 0 else 1,
                            p[pre + 'Conv/%d/bias' % j])
                g += batchnorm(nm + 'BN/%d' % j, nm + 'Conv/%d' % j, p[pre + 'BN/%d/gamma' % j], p[pre + 'BN/%d/beta' % j],
                               p[pre + 'BN/%d/moving_mean' % j], p[pre + 'BN/%d/moving_variance' % j])
                g += leaky_relu(nm + 'LReLU/%d' % j, nm + 'BN/%d' % j)
                nxt.append(nm + 'LReLU/%d' % j)
            cur, cin = nxt, f
            if j == 0 and stride == 2:
                hh, ww = (hh + 1) // 2, (ww + 1) // 2
    n_up = len(net_params['up_conv_kernels'])
    logits_names, softmax_names = [], []
    for t in range(T):
        skips = skips_per_t[t][::-1]
        x, xc = cur[t], cin
        uh, uw = hh, ww
        for ui in range(n_up):
            pre = 'UpLayers/%d/' % ui
            nm = '%st%d/' % (pre, t)
            factor = 2 if ui > 0 else 1
            uh, uw = uh * factor, uw * factor
            g += resize_bilinear(nm + 'resize', x, uh, uw)                       # k.backend.resize_images (:143)
            skip, sc = skips[ui]
            g += concat(nm + 'concat', [nm + 'resize', skip])                    # tf.concat([up, skip]) (:145)
            x, xc = nm + 'concat', xc + sc
            convs = net_params['up_conv_kernels'][ui]
            for j, (k, f) in enumerate(convs):
                g += conv2d(nm + 'Conv/%d' % j, x, p[pre + 'Conv/%d/kernel' % j], 1, p[pre + 'Conv/%d/bias' % j])
                x, xc = nm + 'Conv/%d' % j, f
                if ui == n_up - 1 and j == len(convs) - 1:
                    break                                                         # logits: no BN / activation (:148-149)
                g += batchnorm(nm + 'BN/%d' % j, x, p[pre + 'BN/%d/gamma' % j], p[pre + 'BN/%d/beta' % j],
                               p[pre + 'BN/%d/moving_mean' % j], p[pre + 'BN/%d/moving_variance' % j])
                g += leaky_relu(nm + 'LReLU/%d' % j, nm + 'BN/%d' % j)
                x = nm + 'LReLU/%d' % j
        logits_names.append(x + '/Conv2D')         # OpenCV folds the BiasAdd into the Conv2D layer and keeps that node's name
        g += softmax('softmax_t%d' % t, x)
        softmax_names.append('softmax_t%d' % t)
    return g, {'logits': logits_names, 'softmax': softmax_names, 'states': final_states, 'pad_y': pad_y, 'pad_x': pad_x}


def run_with_opencv(graph_bytes, inp_nhwc, outputs):
    """Execute the graph with OpenCV's TensorFlow importer (plain CPU backend); returns NCHW arrays, one per output name."""
    import cv2
    net = cv2.dnn.readNetFromTensorflow(np.frombuffer(graph_bytes, np.uint8))
    net.setPreferableBackend(cv2.dnn.DNN_BACKEND_OPENCV)
    net.setPreferableTarget(cv2.dnn.DNN_TARGET_CPU)
    net.setInput(np.ascontiguousarray(np.asarray(inp_nhwc, np.float32).transpose(0, 3, 1, 2)))
    outs = net.forward(list(outputs))
    return [np.array(o) for o in outs]


def ulstm_forward_opencv(net_params, params, x_bthwc, pad_image):
    """(logits, softmax) as (B, T, classes, H, W) arrays -- the NCHW API's output layout -- plus the final ConvLSTM states
    [(h, c)] (NCHW) in layer order."""
    x_bthwc = np.asarray(x_bthwc, dtype=np.float32)
    B, T, H, W, C = x_bthwc.shape
    g, names = build_ulstm_graph(net_params, params, B, T, C, H, W, pad_image)
    flat_states = [n for hc in names['states'] for n in hc]
    inp = np.concatenate([x_bthwc[:, t] for t in range(T)], axis=-1)
    outs = run_with_opencv(g, inp, names['logits'] + names['softmax'] + flat_states)
    py, px = names['pad_y'], names['pad_x']

    def crop(a):
        return a[:, :, py[0]:py[0] + H, px[0]:px[0] + W]
    logits = np.stack([crop(o) for o in outs[:T]], axis=1)
    soft = np.stack([crop(o) for o in outs[T:2 * T]], axis=1)
    st = outs[2 * T:]
    return logits, soft, [(st[2 * i], st[2 * i + 1]) for i in range(len(st) // 2)]
