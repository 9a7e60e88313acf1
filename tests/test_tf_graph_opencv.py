"""A third implementation beside the oracle and the CUDA path: the reference network written as the graph of TensorFlow
ops Keras-2 lowers it to (tests/tf_graphdef.py, serialised by hand) and EXECUTED BY OPENCV's TensorFlow importer.

TensorFlow cannot be installed, so this is as close to "the reference's arithmetic run here" as the image
allows: OpenCV's implementation of Conv2D / SAME, FusedBatchNormV3, LeakyRelu, ResizeBilinear (half-pixel), MirrorPad,
ConcatV2, Softmax and the element-wise ops of the ConvLSTM cell shares no code with torch (the oracle) or with this
repository.  Per operator, per ConvLSTM step and for whole networks the oracle must agree with it to float32 rounding; the
committed vectors it produced (tests/golden/tf_graph_opencv.npz) are then the target of the host build here and of the
tcgen05 path on the GPU (north_star tolerance: 1e-3 relative in the bf16x3 parity mode)."""
import os

import numpy as np
import pytest
import torch

from oracle import lstm_unet_oracle as O
from oracle import np_semantics as S
from tests import tf_graphdef as G
from tests.golden.make_golden import NET
from tests.golden.make_tf_graph_golden import NET_ODD

cv2 = pytest.importorskip('cv2')
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

NET_WIDE = {   # CTC channel structure at 1/4 width (every buffer a multiple of 64 except the 32 / 16-wide tail)
    'down_conv_kernels': [[(3, 64), (3, 64)], [(3, 128), (3, 128)], [(3, 128), (3, 128)], [(3, 192), (3, 192)]],
    'lstm_kernels': [[(5, 64)], [(5, 128)], [(5, 128)], [(5, 192)]],
    'up_conv_kernels': [[(3, 128), (3, 128)], [(3, 64), (3, 64)], [(3, 32), (3, 32)], [(3, 16), (3, 16), (1, 3)]],
}
NET_TWO = {
    'down_conv_kernels': [[(3, 6)], [(3, 10), (3, 10)]],
    'lstm_kernels': [[(3, 5), (3, 7)], [(5, 9)]],
    'up_conv_kernels': [[(3, 6)], [(3, 5), (1, 3)]],
}


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def nchw(a):
    return np.ascontiguousarray(np.asarray(a).transpose(0, 3, 1, 2))


def run1(graph, x_nhwc, out):
    return G.run_with_opencv(graph, x_nhwc, [out])[0]


# ---- operator by operator -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("H,W,k,s,cin,cout", [
    (10, 12, 3, 2, 3, 5),      # even sizes, stride 2: SAME pads 0 before / 1 after
    (11, 13, 3, 2, 2, 4),      # odd sizes, stride 2: SAME pads 1 / 1
    (8, 7, 3, 2, 1, 2),        # mixed
    (9, 9, 5, 1, 2, 3), (7, 10, 3, 1, 4, 4), (6, 5, 1, 1, 3, 2), (12, 12, 5, 2, 2, 2),
])
def test_conv2d_same_as_opencv_runs_it(H, W, k, s, cin, cout):
    rng = np.random.default_rng(H * 100 + W)
    x = rng.standard_normal((2, H, W, cin)).astype(np.float32)
    w = rng.standard_normal((k, k, cin, cout)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    g = G.placeholder('x', [2, H, W, cin]) + G.conv2d('y', 'x', w, s, b)
    got = run1(g, x, 'y/Conv2D')
    ora = O.conv2d_same(torch.from_numpy(nchw(x)), torch.from_numpy(w), torch.from_numpy(b), s).numpy()
    assert got.shape == ora.shape == (2, cout, -(-H // s), -(-W // s))
    assert rel(got, ora) < 1e-5
    assert rel(got, nchw(S.conv2d_same_nhwc(x, w, b, s))) < 1e-5


def test_batchnorm_leakyrelu_resize_mirrorpad_softmax_as_opencv_runs_them():
    rng = np.random.default_rng(3)
    H, W, C = 6, 7, 4
    x = rng.standard_normal((1, H, W, C)).astype(np.float32)
    xt = torch.from_numpy(nchw(x))
    gam, bet, mean = [rng.standard_normal(C).astype(np.float32) for _ in range(3)]
    var = rng.uniform(1e-4, 2, C).astype(np.float32)               # small variances: where epsilon = 1e-3 enters matters
    ph = G.placeholder('x', [1, H, W, C])
    got = run1(ph + G.batchnorm('bn', 'x', gam, bet, mean, var), x, 'bn')
    ora = O.batchnorm(xt, *[torch.from_numpy(a) for a in (gam, bet, mean, var)], False)
    ora = ora[0] if isinstance(ora, tuple) else ora
    assert rel(got, ora.numpy()) < 1e-5
    assert np.array_equal(run1(ph + G.leaky_relu('a', 'x'), x, 'a'), O.leaky_relu(xt).numpy())      # alpha = 0.3
    got = run1(ph + G.resize_bilinear('up', 'x', 2 * H, 2 * W), x, 'up')
    assert rel(got, O.resize_bilinear(xt, 2).numpy()) < 1e-6 and rel(got, nchw(S.bilinear_up_nhwc(x, 2))) < 1e-6
    assert rel(run1(ph + G.resize_bilinear('same', 'x', H, W), x, 'same'), nchw(x)) < 1e-7          # the factor-1 UpBlock
    got = run1(ph + G.mirror_pad('p', 'x', 2, 3, 1, 4), x, 'p')
    assert np.array_equal(got, S.reflect_pad_hw(nchw(x), 2, 3, 1, 4))
    assert rel(run1(ph + G.softmax('s', 'x'), x, 's'), torch.softmax(xt, 1).numpy()) < 1e-6


@pytest.mark.parametrize("k,scale", [(5, 0.2), (3, 1.5)])          # scale 1.5 saturates the hard-sigmoid on both sides
def test_convlstm_cell_as_opencv_runs_it(k, scale):
    rng = np.random.default_rng(k)
    H, W, Cin, F_ = 6, 7, 3, 4
    x = rng.standard_normal((1, H, W, Cin)).astype(np.float32)
    h0 = (rng.standard_normal((1, H, W, F_)) * 0.5).astype(np.float32)
    c0 = (rng.standard_normal((1, H, W, F_)) * 0.5).astype(np.float32)
    wk = (rng.standard_normal((k, k, Cin, 4 * F_)) * scale).astype(np.float32)
    wr = (rng.standard_normal((k, k, F_, 4 * F_)) * scale).astype(np.float32)
    b = rng.standard_normal(4 * F_).astype(np.float32)
    inp = np.concatenate([x, h0, c0], axis=3)
    ct = inp.shape[3]
    g = G.placeholder('inp', [1, H, W, ct]) + G.select_channels('x', 'inp', ct, 0, Cin)
    g += G.select_channels('h0', 'inp', ct, Cin, F_) + G.select_channels('c0', 'inp', ct, Cin + F_, F_)
    gg, hn, cn = G.convlstm_cell('cell', 'x', 'h0', 'c0', wk, wr, b)
    h1, c1 = G.run_with_opencv(g + gg, inp, [hn, cn])
    rh, rc = S.convlstm_step_nhwc(x, h0, c0, wk, wr, b)
    z = S.conv2d_same_nhwc(x, wk, b, 1) + S.conv2d_same_nhwc(h0, wr, None, 1)
    if scale > 1:
        lin = 0.2 * z[..., :F_] + 0.5
        assert (lin < 0).any() and (lin > 1).any()                  # both clip branches are exercised
    assert rel(h1, nchw(rh)) < 5e-5 and rel(c1, nchw(rc)) < 5e-5


# ---- whole networks ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("net,B,T,H,W,pad", [
    (NET_TWO, 1, 2, 18, 22, False),        # two ConvLSTM layers in one level
    (NET_TWO, 2, 2, 17, 21, True),         # odd sizes + pad_image
    (NET_ODD, 1, 2, 35, 35, True),         # the reference unit_test's shape (Networks.py:256-277)
    (NET_ODD, 2, 3, 40, 48, True),
    (NET, 2, 3, 24, 32, False),
    (NET_WIDE, 1, 2, 64, 48, False),       # the wide case of test_gpu_forward's parity list
])
def test_oracle_matches_the_tf_graph_run_by_opencv(net, B, T, H, W, pad):
    params = O.init_params(net, seed=3, randomize_bn=True)
    x = np.random.default_rng(0).standard_normal((B, T, 1, H, W)).astype(np.float32)
    ora = O.OracleNet(net, 'NCHW', pad, params=params)
    ref_l, ref_s = ora(torch.from_numpy(x), False)
    logits, soft, states = G.ulstm_forward_opencv(net, {k: v.numpy() for k, v in params.items()},
                                                  x.transpose(0, 1, 3, 4, 2), pad)
    assert logits.shape == tuple(ref_l.shape) == (B, T, 3, H, W)
    assert rel(logits, ref_l.numpy()) < 5e-5 and rel(soft, ref_s.numpy()) < 5e-5
    flat = [hc for lvl in ora.get_states() for hc in lvl]
    assert len(flat) == len(states)
    for (h, c), (rh, rc) in zip(states, flat):
        assert rel(h, rh) < 5e-5 and rel(c, rc) < 5e-5


@pytest.mark.skipif(not os.path.exists('/root/reference/Networks.py'), reason='the reference is only present in the build container')
@pytest.mark.parametrize("net,B,T,H,W,pad,calls", [
    (NET_TWO, 1, 2, 18, 22, False, 2), (NET_TWO, 2, 1, 17, 21, True, 3), (NET_ODD, 1, 2, 35, 35, True, 1),
])
def test_graph_emitted_by_the_references_own_networks_py(net, B, T, H, W, pad, calls):
    """The reference's unmodified ULSTMnet2D.call run on the graph-emitting stand-in (tests/keras_graph_standin.py): its
    graph, executed by OpenCV, gives bit for bit what the reference-free builder's graph gives (so the committed vectors are
    'the reference's wiring + OpenCV's arithmetic'), and the oracle agrees with both, stateful calls included."""
    from tests.golden.make_tf_graph_golden import reference_driven
    params = O.init_params(net, seed=21, randomize_bn=True)
    pn = {k: v.numpy() for k, v in params.items()}
    rng = np.random.default_rng(1)
    xs = [rng.standard_normal((B, T, 1, H, W)).astype(np.float32) for _ in range(calls)]
    per_call, states = reference_driven(net, pn, xs, pad)
    seq = np.concatenate(xs, axis=1)
    logits, soft, st = G.ulstm_forward_opencv(net, pn, seq.transpose(0, 1, 3, 4, 2), pad)
    ora = O.OracleNet(net, 'NCHW', pad, params=params)
    for c, (rl, rs) in enumerate(per_call):
        assert np.array_equal(rl, logits[:, c * T:(c + 1) * T]) and np.array_equal(rs, soft[:, c * T:(c + 1) * T])
        ol, os_ = ora(torch.from_numpy(xs[c]), False)
        assert rel(rl, ol.numpy()) < 5e-5 and rel(rs, os_.numpy()) < 5e-5
    for (h, c_), (h2, c2) in zip(states, st):
        assert np.array_equal(h, h2) and np.array_equal(c_, c2)


def load_gold():
    return np.load(os.path.join(GOLD, 'tf_graph_opencv.npz'))


def test_committed_vectors_are_what_opencv_produces_and_agree_with_the_fp64_oracle_vectors():
    z = load_gold()
    f = np.load(os.path.join(GOLD, 'forward_pad.npz'))
    # the fp64-oracle vectors of the same case: two stateful calls of 2 frames == frames 0-1 / 2-3 of the unrolled graph
    assert rel(z['pad:logits'][:, 0:2], f['logits0']) < 5e-5 and rel(z['pad:logits'][:, 2:4], f['logits1']) < 5e-5
    assert rel(z['pad:softmax'][:, 2:4], f['softmax1']) < 5e-5
    assert rel(z['pad:h_lvl0'], f['h_lvl0']) < 5e-5 and rel(z['pad:c_lvl2'], f['c_lvl2']) < 5e-5
    # regenerate the second case: the file is OpenCV's output, not an edited copy
    p = {k[len('odd:p:'):]: z[k] for k in z.files if k.startswith('odd:p:')}
    logits, soft, states = G.ulstm_forward_opencv(NET_ODD, p, z['odd:x'].transpose(0, 1, 3, 4, 2), True)
    # (bit-identical on the machine that wrote the file; another CPU's SIMD path may sum in another order)
    assert rel(logits, z['odd:logits']) < 1e-5 and rel(soft, z['odd:softmax']) < 1e-5
    assert rel(states[3][1], z['odd:c3']) < 1e-5


def test_host_build_matches_the_opencv_vectors():
    from tests.emu_backend import emu_session, emu_forward
    z = load_gold()
    p = {k[len('odd:p:'):]: z[k] for k in z.files if k.startswith('odd:p:')}
    sess = emu_session(NET_ODD, data_format='NCHW', pad_image=True, batch=2, max_t=3, height=35, width=35, precision='bf16x3')
    sess.set_params(p)
    logits, softmax = emu_forward(sess, z['odd:x'], False)
    assert rel(logits, z['odd:logits']) < 1e-3 and rel(softmax, z['odd:softmax']) < 1e-3
    for lvl in range(4):
        out = np.zeros(sess.state_shape(lvl, 0), np.float32)
        sess.get_state(lvl, 0, 1, out.ctypes.data)
        assert rel(out, z['odd:c%d' % lvl]) < 1e-3, lvl
    sess.close()
    # the throughput mode (bf16 operands, 8 significant bits; measured here: 6e-3 / 2e-3)
    sess = emu_session(NET_ODD, data_format='NCHW', pad_image=True, batch=2, max_t=3, height=35, width=35, precision='bf16')
    sess.set_params(p)
    logits, softmax = emu_forward(sess, z['odd:x'], False)
    assert rel(logits, z['odd:logits']) < 2e-2 and rel(softmax, z['odd:softmax']) < 2e-2
    sess.close()


@pytest.mark.gpu
def test_tcgen05_matches_the_tf_graph_run_by_opencv():
    """The CUDA path against vectors that neither the oracle nor this repository produced."""
    from lstm_unet_b200.Networks import ULSTMnet2D
    z = load_gold()
    p = {k[len('odd:p:'):]: z[k] for k in z.files if k.startswith('odd:p:')}
    m = ULSTMnet2D(NET_ODD, 'NCHW', True, precision='bf16x3')
    m.set_weights_dict(p)
    logits, softmax = m(z['odd:x'], False)
    assert rel(logits.numpy(), z['odd:logits']) < 1e-3 and rel(softmax.numpy(), z['odd:softmax']) < 1e-3
    st = m.get_states()
    for lvl in range(4):
        assert rel(st[lvl][0][0], z['odd:h%d' % lvl]) < 1e-3 and rel(st[lvl][0][1], z['odd:c%d' % lvl]) < 1e-3, lvl
    m = ULSTMnet2D(NET_ODD, 'NCHW', True, precision='bf16')            # the bench's precision: 5e-2 as in test_gpu_forward
    m.set_weights_dict(p)
    logits, softmax = m(z['odd:x'], False)
    assert rel(logits.numpy(), z['odd:logits']) < 5e-2 and rel(softmax.numpy(), z['odd:softmax']) < 5e-2
    f = np.load(os.path.join(GOLD, 'forward_pad.npz'))
    m = ULSTMnet2D(NET, 'NCHW', True, precision='bf16x3')
    m.set_weights_dict({k[2:]: f[k] for k in f.files if k.startswith('p:')})
    for call in range(2):                                             # stateful carry == frames 2-3 of the unrolled graph
        logits, _ = m(f['x'][call], False)
        assert rel(logits.numpy(), z['pad:logits'][:, 2 * call:2 * call + 2]) < 1e-3, call
    assert m.launch_count() > 0


def test_hand_written_graphdef_parses_with_tensorflows_own_proto_schema():
    """TensorBoard ships TensorFlow's compiled .proto definitions: the bytes tf_graphdef.py writes must decode with them into
    the nodes, attributes and tensors that were meant (field numbers, wire types, packed lists, tensor_content)."""
    graph_pb2 = pytest.importorskip('tensorboard.compat.proto.graph_pb2')
    rng = np.random.default_rng(0)
    w = rng.standard_normal((3, 3, 2, 4)).astype(np.float32)
    b = rng.standard_normal(4).astype(np.float32)
    raw = G.placeholder('x', [1, 8, 9, 2]) + G.conv2d('c', 'x', w, 2, b) + G.batchnorm('bn', 'c', b, b, b, np.abs(b)) + \
        G.leaky_relu('a', 'bn') + G.resize_bilinear('up', 'a', 8, 10) + G.concat('cat', ['up', 'up']) + \
        G.mirror_pad('pad', 'cat', 1, 2, 3, 4) + G.softmax('sm', 'pad')
    gd = graph_pb2.GraphDef.FromString(raw)
    nodes = {n.name: n for n in gd.node}
    assert [n.op for n in gd.node if n.op != 'Const'] == ['Placeholder', 'Conv2D', 'BiasAdd', 'FusedBatchNormV3', 'LeakyRelu',
                                                          'ResizeBilinear', 'ConcatV2', 'MirrorPad', 'Softmax']
    ph = nodes['x']
    assert ph.attr['dtype'].type == 1 and [d.size for d in ph.attr['shape'].shape.dim] == [1, 8, 9, 2]
    conv = nodes['c/Conv2D']
    assert list(conv.input) == ['x', 'c/kernel'] and conv.attr['padding'].s == b'SAME' and conv.attr['data_format'].s == b'NHWC'
    assert list(conv.attr['strides'].list.i) == [1, 2, 2, 1] and conv.attr['T'].type == 1
    k = nodes['c/kernel'].attr['value'].tensor
    assert k.dtype == 1 and [d.size for d in k.tensor_shape.dim] == [3, 3, 2, 4]
    assert np.array_equal(np.frombuffer(k.tensor_content, np.float32).reshape(3, 3, 2, 4), w)
    assert list(nodes['c'].input) == ['c/Conv2D', 'c/bias']
    bn = nodes['bn']
    assert abs(bn.attr['epsilon'].f - 1e-3) < 1e-9 and bn.attr['is_training'].b is False and len(bn.input) == 5
    assert abs(nodes['a'].attr['alpha'].f - 0.3) < 1e-7
    up = nodes['up']
    assert up.attr['half_pixel_centers'].b is True and up.attr['align_corners'].b is False
    size = nodes['up/size'].attr['value'].tensor
    assert size.dtype == 3 and list(np.frombuffer(size.tensor_content, np.int32)) == [8, 10]
    axis = nodes['cat/axis'].attr['value'].tensor
    assert axis.dtype == 3 and list(axis.int_val) == [3] and len(axis.tensor_shape.dim) == 0 and nodes['cat'].attr['N'].i == 2
    pads = nodes['pad/paddings'].attr['value'].tensor
    assert np.frombuffer(pads.tensor_content, np.int32).reshape(4, 2).tolist() == [[0, 0], [1, 2], [3, 4], [0, 0]]
    assert nodes['pad'].attr['mode'].s == b'REFLECT'
    # and a whole network: every input of every node names a node of the graph, no name is used twice
    params = {k_: v.numpy() for k_, v in O.init_params(NET_TWO, seed=1, randomize_bn=True).items()}
    g, names = G.build_ulstm_graph(NET_TWO, params, 1, 2, 1, 10, 12, True)
    gd = graph_pb2.GraphDef.FromString(g)
    seen = set()
    for n in gd.node:
        assert n.name not in seen, n.name
        assert all(i in seen for i in n.input), (n.name, list(n.input))       # topological order, as TensorFlow writes it
        seen.add(n.name)
    assert all(x in seen for x in names['softmax']) and all(h in seen and c in seen for h, c in names['states'])
    assert sum(n.op == 'Conv2D' for n in gd.node) > 50


@pytest.mark.gpu
def test_tcgen05_wide_network_matches_the_tf_graph_run_by_opencv_live():
    """Multiple-of-64 channel counts (full tensor-core tiles, two-source decoder convs) against OpenCV run on the spot: the
    weights are too large to commit as vectors, and the reference-free graph builder is bit-identical to the reference-driven
    one (test_graph_emitted_by_the_references_own_networks_py)."""
    from lstm_unet_b200.Networks import ULSTMnet2D
    params = {k: v.numpy() for k, v in O.init_params(NET_WIDE, seed=7, randomize_bn=True).items()}
    x = np.random.default_rng(2).standard_normal((1, 2, 1, 64, 48)).astype(np.float32)
    ref_l, ref_s, ref_st = G.ulstm_forward_opencv(NET_WIDE, params, x.transpose(0, 1, 3, 4, 2), False)
    m = ULSTMnet2D(NET_WIDE, 'NCHW', False, precision='bf16x3')
    m.set_weights_dict({k: v.copy() for k, v in params.items()})
    logits, softmax = m(x, False)
    assert rel(logits.numpy(), ref_l) < 1e-3 and rel(softmax.numpy(), ref_s) < 1e-3
    st = m.get_states()
    for lvl in range(4):
        assert rel(st[lvl][0][0], ref_st[lvl][0]) < 1e-3 and rel(st[lvl][0][1], ref_st[lvl][1]) < 1e-3, lvl
