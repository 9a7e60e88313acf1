"""Independent loop-level numpy restatement of the operator semantics -- TEST INFRASTRUCTURE ONLY.

Written from the published TensorFlow-2 / Keras-2 operator definitions, with explicit
index formulas instead of library calls, so that ``oracle/lstm_unet_oracle.py`` (torch library calls) is
pinned against a second, structurally different statement of the same semantics.  fp64, NHWC, slow: small
shapes only.  (The reference itself cannot run here: TensorFlow is absent -- "parity unpinned".)
"""
import numpy as np


def same_pads(n, k, s):
    out = (n + s - 1) // s
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2


def conv2d_same_nhwc(x, w, b, stride):
    """x (N,H,W,Ci), w (kh,kw,Ci,Co) HWIO, cross-correlation: out[y,x] = sum w[r,s] * in[y*st+r-pt, x*st+s-pl]."""
    N, H, W, Ci = x.shape
    kh, kw, _, Co = w.shape
    Ho, pt = same_pads(H, kh, stride)
    Wo, pl = same_pads(W, kw, stride)
    y = np.zeros((N, Ho, Wo, Co), dtype=np.float64)
    for oy in range(Ho):
        for ox in range(Wo):
            acc = np.zeros((N, Co), dtype=np.float64)
            for r in range(kh):
                iy = oy * stride + r - pt
                if iy < 0 or iy >= H:
                    continue
                for s in range(kw):
                    ix = ox * stride + s - pl
                    if ix < 0 or ix >= W:
                        continue
                    acc += x[:, iy, ix, :] @ w[r, s]
            y[:, oy, ox, :] = acc
    if b is not None:
        y += b
    return y


def hard_sigmoid(x):
    return np.minimum(np.maximum(0.2 * x + 0.5, 0.0), 1.0)


def convlstm_step_nhwc(x_t, h, c, wk, wr, b):
    F_ = wr.shape[2]
    z = conv2d_same_nhwc(x_t, wk, b, 1) + conv2d_same_nhwc(h, wr, None, 1)
    i = hard_sigmoid(z[..., 0:F_])
    f = hard_sigmoid(z[..., F_:2 * F_])
    g = np.tanh(z[..., 2 * F_:3 * F_])
    o = hard_sigmoid(z[..., 3 * F_:4 * F_])
    c2 = f * c + i * g
    h2 = o * np.tanh(c2)
    return h2, c2


def bilinear_up_nhwc(x, f):
    """tf.image.resize bilinear, half_pixel_centers=True, no antialias, integer factor f."""
    N, H, W, C = x.shape
    out = np.zeros((N, H * f, W * f, C), dtype=np.float64)
    for oy in range(H * f):
        sy = (oy + 0.5) / f - 0.5
        y0 = int(np.floor(sy))
        wy = sy - y0
        ya, yb = min(max(y0, 0), H - 1), min(max(y0 + 1, 0), H - 1)
        for ox in range(W * f):
            sx = (ox + 0.5) / f - 0.5
            x0 = int(np.floor(sx))
            wx = sx - x0
            xa, xb = min(max(x0, 0), W - 1), min(max(x0 + 1, 0), W - 1)
            top = x[:, ya, xa] * (1 - wx) + x[:, ya, xb] * wx
            bot = x[:, yb, xa] * (1 - wx) + x[:, yb, xb] * wx
            out[:, oy, ox] = top * (1 - wy) + bot * wy
    return out


def batchnorm_train_nhwc(x, gamma, beta, eps=1e-3):
    n = x.shape[0] * x.shape[1] * x.shape[2]
    mean = x.reshape(n, -1).sum(0) / n
    var = ((x.reshape(n, -1) - mean) ** 2).sum(0) / n
    return (x - mean) / np.sqrt(var + eps) * gamma + beta, mean, var * n / max(n - 1, 1)


def leaky_relu(x, alpha=0.3):
    return np.where(x > 0, x, alpha * x)


def reflect_pad_hw(x, pt, pb, pl, pr):
    """tf.pad(..., 'REFLECT') on an (...,H,W) array: mirror without repeating the edge."""
    H, W = x.shape[-2], x.shape[-1]
    ys = [abs(i) if i < H else 2 * (H - 1) - i for i in range(-pt, H + pb)]
    xs = [abs(i) if i < W else 2 * (W - 1) - i for i in range(-pl, W + pr)]
    return x[..., ys, :][..., xs]


def weighted_ce(labels, logits, class_weights):
    """labels (...,) in {-1,0,1,2}; logits (...,3).  losses.py:13-27."""
    lab = labels.astype(np.int64)
    valid = (labels > -1).astype(np.float64)
    tot = 0.0
    flat_l, flat_z, flat_v = lab.reshape(-1), logits.reshape(-1, 3), valid.reshape(-1)
    for i in range(flat_l.shape[0]):
        if flat_l[i] < 0:
            continue                      # one_hot(-1) = 0 and valid = 0
        z = flat_z[i]
        m = z.max()
        lse = m + np.log(np.exp(z - m).sum())
        tot += (lse - z[flat_l[i]]) * class_weights[flat_l[i]] * flat_v[i]
    return tot / (valid.sum() + 0.00001)
