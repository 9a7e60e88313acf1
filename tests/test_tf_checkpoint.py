"""TF2 checkpoint (tensor bundle) reader / writer without TensorFlow: known-answer vectors for CRC-32C and Snappy,
SSTable round trips (multi-block, prefix compression, a Snappy-compressed block), bundle round trip with the Keras
object-graph keys of ULSTMnet2D."""
import os
import struct

import numpy as np
import pytest

from lstm_unet_b200 import tf_checkpoint as T


def test_crc32c_known_answers():
    assert T.crc32c(b'123456789') == 0xE3069283            # the standard CRC-32C check value
    assert T.crc32c(b'') == 0
    assert T.crc32c(bytes(32)) == 0x8A9136AA               # RFC 3720 B.4: 32 bytes of zeros
    assert T.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43      # RFC 3720 B.4: 32 bytes of ones
    assert T.mask_crc(0) == 0xa282ead8


def test_snappy_decoder_vectors():
    # literal only
    assert T.snappy_uncompress(bytes([5, (5 - 1) << 2]) + b'hello') == b'hello'
    # literal "ab" + copy (1-byte offset form): len 6, offset 2  -> "abababab"
    comp = bytes([8, (2 - 1) << 2]) + b'ab' + bytes([((6 - 4) << 2) | 1, 2])
    assert T.snappy_uncompress(comp) == b'abababab'
    # 2-byte offset copy: literal "0123456789" then copy len 10 offset 10
    comp = bytes([20, (10 - 1) << 2]) + b'0123456789' + bytes([((10 - 1) << 2) | 2, 10, 0])
    assert T.snappy_uncompress(comp) == b'01234567890123456789'
    # long literal (length in one extra byte)
    lit = bytes(range(200))
    comp = bytes([200, 1, 60 << 2, 199]) + lit
    assert T.snappy_uncompress(comp) == lit
    with pytest.raises(ValueError):
        T.snappy_uncompress(bytes([4, ((4 - 4) << 2) | 1, 9]))      # copy before any output


def test_sstable_round_trip_multi_block(tmp_path):
    items = [(('key/%05d/suffix' % i).encode(), os.urandom(1 + i % 97)) for i in range(700)]
    items.sort()
    path = str(tmp_path / 't.index')
    T.write_table(path, items, block_size=512)
    got = T.read_table(path)
    assert got == dict(items)
    raw = bytearray(open(path, 'rb').read())
    raw[10] ^= 0xFF                                          # corrupt a data block -> checksum error
    open(path, 'wb').write(bytes(raw))
    with pytest.raises(ValueError):
        T.read_table(path)


def test_reads_snappy_compressed_block(tmp_path):
    """A table whose data block is stored Snappy-compressed (type byte 1), as TensorFlow's table builder does when it pays."""
    items = [(b'a/%d' % i, b'v' * 20) for i in range(5)]
    block = T._build_block(items)
    # all-literal snappy encoding of the block
    n = len(block)
    comp = T._put_varint(n) + bytes([60 << 2, n - 1]) + block if n > 60 else T._put_varint(n) + bytes([(n - 1) << 2]) + block
    f = bytearray()
    f += comp + b'\x01' + struct.pack('<I', T.mask_crc(T.crc32c(comp + b'\x01')))
    data_handle = T._put_varint(0) + T._put_varint(len(comp))

    def emit(blk):
        off = len(f)
        f.extend(blk + b'\x00' + struct.pack('<I', T.mask_crc(T.crc32c(blk + b'\x00'))))
        return off, len(blk)
    moff, msize = emit(T._build_block([]))
    ioff, isize = emit(T._build_block([(items[-1][0], data_handle)], restart_interval=1))
    footer = T._put_varint(moff) + T._put_varint(msize) + T._put_varint(ioff) + T._put_varint(isize)
    f += footer + b'\x00' * (40 - len(footer)) + struct.pack('<Q', T.TABLE_MAGIC)
    path = str(tmp_path / 's.index')
    open(path, 'wb').write(bytes(f))
    assert T.read_table(path) == dict(items)


def test_bundle_round_trip_with_keras_keys(tmp_path):
    from oracle import lstm_unet_oracle as O
    net = {'down_conv_kernels': [[(3, 6)], [(3, 10), (3, 10)]], 'lstm_kernels': [[(3, 5), (3, 7)], [(5, 9)]],
           'up_conv_kernels': [[(3, 6)], [(3, 5), (1, 3)]]}
    params = {k: v.numpy() for k, v in O.init_params(net, seed=3, randomize_bn=True).items()}
    prefix = str(tmp_path / 'model.ckpt')
    T.save_model_weights(prefix, params)
    assert os.path.exists(prefix + '.index') and os.path.exists(prefix + '.data-00000-of-00001')
    raw = T.read_bundle(prefix)
    assert 'DownLayers/0/ConvLSTM/1/cell/recurrent_kernel/.ATTRIBUTES/VARIABLE_VALUE' in raw
    assert 'UpLayers/1/Conv/1/bias/.ATTRIBUTES/VARIABLE_VALUE' in raw
    assert 'DownLayers/1/BN/0/moving_variance/.ATTRIBUTES/VARIABLE_VALUE' in raw
    got = T.load_model_weights(prefix, list(params))
    for k in params:
        np.testing.assert_array_equal(got[k], params[k])
    # the same variables under tf.train.Checkpoint(net=model) (train2D.py:62)
    T.save_model_weights(prefix + '2', params, root='net/')
    got = T.load_model_weights(prefix + '2', list(params))
    np.testing.assert_array_equal(got['DownLayers/0/Conv/0/kernel'], params['DownLayers/0/Conv/0/kernel'])
    with pytest.raises(KeyError):
        T.load_model_weights(prefix, list(params) + ['DownLayers/9/Conv/0/kernel'])


def test_model_save_load_tf_format(tmp_path):
    """ULSTMnet2D.save_weights(save_format='tf') / load_weights(prefix) -- no GPU needed before the first call."""
    from lstm_unet_b200.Networks import ULSTMnet2D
    from oracle import lstm_unet_oracle as O
    net = {'down_conv_kernels': [[(3, 6)]], 'lstm_kernels': [[(3, 5)]], 'up_conv_kernels': [[(3, 5), (1, 3)]]}
    params = {k: v.numpy() for k, v in O.init_params(net, seed=4, randomize_bn=True).items()}
    m = ULSTMnet2D(net, 'NCHW', True)
    m.set_weights_dict(params)
    m.save_weights(str(tmp_path / 'model.ckpt'), save_format='tf')
    m2 = ULSTMnet2D(net, 'NCHW', True)
    m2.load_weights(str(tmp_path / 'model.ckpt'))
    for k in params:
        np.testing.assert_array_equal(m2._pending_weights[k], params[k])


def test_training_checkpoint_manager_round_trip(tmp_path):
    """checkpoint.Checkpoint / CheckpointManager / latest_checkpoint (train2D.py:62-85,222-226) on the host: model
    variables under net/, Adam moments as optimizer slots, int64 counters, max_to_keep, the `checkpoint` state file"""
    from lstm_unet_b200 import checkpoint as ck
    from lstm_unet_b200 import tf_checkpoint as tfc
    from lstm_unet_b200.Networks import ULSTMnet2D, Adam
    from oracle import lstm_unet_oracle as O
    net = {'down_conv_kernels': [[(3, 6)], [(3, 8)]], 'lstm_kernels': [[(3, 5)], [(5, 7)]], 'up_conv_kernels': [[(3, 6)], [(3, 5), (1, 3)]]}
    params = {k: v.numpy() for k, v in O.init_params(net, seed=2, randomize_bn=True).items()}
    model = ULSTMnet2D(net, 'NCHW', False, train=True)
    model.set_weights_dict(params)
    layout = model._variable_layout()
    n_train = sum(e['count'] for e in layout if e['trainable'])
    rng = np.random.default_rng(0)
    opt = Adam(lr=3e-4)
    opt.set_slots(17, rng.standard_normal(n_train).astype(np.float32), rng.random(n_train).astype(np.float32))
    c = ck.Checkpoint(model, opt, step=17)
    mgr = ck.CheckpointManager(c, str(tmp_path / 'tf_ckpts'), max_to_keep=2)
    assert ck.latest_checkpoint(str(tmp_path / 'tf_ckpts')) is None
    paths = []
    for step in (17, 18, 19):
        c.step = step
        paths.append(mgr.save(step))
    assert [os.path.basename(p) for p in mgr.checkpoints] == ['ckpt-18', 'ckpt-19']
    assert not os.path.exists(paths[0] + '.index') and os.path.exists(paths[2] + '.data-00000-of-00001')
    assert ck.latest_checkpoint(str(tmp_path / 'tf_ckpts')) == paths[2]
    bundle = tfc.read_bundle(paths[2])
    assert bundle['step/.ATTRIBUTES/VARIABLE_VALUE'].dtype == np.int64 and int(bundle['step/.ATTRIBUTES/VARIABLE_VALUE'].reshape(-1)[0]) == 19
    assert 'net/DownLayers/0/ConvLSTM/0/cell/kernel/.OPTIMIZER_SLOT/optimizer/m/.ATTRIBUTES/VARIABLE_VALUE' in bundle
    # restore into a fresh model / optimizer
    model2, opt2 = ULSTMnet2D(net, 'NCHW', False, train=True), Adam(lr=1.0)
    c2 = ck.Checkpoint(model2, opt2).restore(ck.latest_checkpoint(str(tmp_path / 'tf_ckpts')))
    assert c2.step == 19 and opt2.iterations == 17
    for k, v in params.items():
        assert np.array_equal(model2._pending_weights[k], v), k
    it, m, v = opt.get_slots()
    it2, m2, v2 = opt2.get_slots()
    assert np.array_equal(m, m2) and np.array_equal(v, v2)
    # a manager opened on an existing directory continues the list; a plain save_weights file restores weights only
    mgr2 = ck.CheckpointManager(c2, str(tmp_path / 'tf_ckpts'), max_to_keep=2)
    assert mgr2.latest_checkpoint == paths[2]
    model.save_weights(str(tmp_path / 'model.ckpt'), save_format='tf')
    opt3 = Adam()
    c3 = ck.Checkpoint(ULSTMnet2D(net, 'NCHW', False), opt3).restore(str(tmp_path / 'model.ckpt'))
    assert c3.step == 0 and opt3.iterations == 0 and opt3.get_slots()[1] is None


def test_crc32c_and_masking_against_tensorboards_implementation():
    """TensorBoard carries its own CRC-32C + TensorFlow masking (for TFRecord event files): a third-party implementation of
    the checksum every table block and every tensor of a bundle is guarded with."""
    pw = pytest.importorskip('tensorboard.compat.tensorflow_stub.pywrap_tensorflow')
    rng = np.random.default_rng(0)
    for n in (0, 1, 3, 4, 5, 63, 64, 65, 1000, 4099):
        buf = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        assert T.crc32c(buf) == pw.crc32c(buf), n
        assert T.mask_crc(T.crc32c(buf)) == pw.masked_crc32c(buf), n
    a, b = os.urandom(100), os.urandom(77)
    assert T.crc32c(b, T.crc32c(a)) == pw.crc32c(a + b)          # incremental form used for the multi-part tensors


def test_object_graph_is_written_and_parses_with_the_tensorflow_schema(tmp_path):
    """``save_weights(save_format='tf')`` and the training checkpoints carry ``_CHECKPOINTABLE_OBJECT_GRAPH``: a scalar string
    tensor (lengths + length checksum + bytes) holding a TrackableObjectGraph.  Parsed here with the TensorFlow proto schema
    that TensorBoard bundles: every variable is reachable from the root along its Keras attribute path
    (DownLayers -> 0 -> ConvLSTM -> 0 -> cell -> kernel), carries its checkpoint key, and the Adam moments hang off the
    optimizer node as slot variables of the right original variable."""
    pb = pytest.importorskip('tensorboard.compat.proto.trackable_object_graph_pb2')
    from lstm_unet_b200 import checkpoint as ck
    from lstm_unet_b200.Networks import ULSTMnet2D, Adam
    from oracle import lstm_unet_oracle as O
    net = {'down_conv_kernels': [[(3, 6)]], 'lstm_kernels': [[(3, 5)]], 'up_conv_kernels': [[(3, 5), (1, 3)]]}
    params = {k: v.numpy() for k, v in O.init_params(net, seed=4, randomize_bn=True).items()}
    m = ULSTMnet2D(net, 'NCHW', True, train=True)
    m.set_weights_dict(params)
    prefix = str(tmp_path / 'model.ckpt')
    m.save_weights(prefix, save_format='tf')
    raw = T.read_bundle(prefix, strings=True)
    g = pb.TrackableObjectGraph()
    g.ParseFromString(raw[T.OBJECT_GRAPH_KEY])

    def follow(graph, path):
        nid = 0
        for part in path.split('/'):
            nxt = [c.node_id for c in graph.nodes[nid].children if c.local_name == part]
            assert len(nxt) == 1, (path, part)
            nid = nxt[0]
        return nid
    for name in params:
        key = T.keras_key(name)
        node = g.nodes[follow(g, key[:-len(T.VAR_SUFFIX)])]
        assert [(a.name, a.checkpoint_key) for a in node.attributes] == [('VARIABLE_VALUE', key)]
        assert key in raw
    # training checkpoint: net/ prefix, step, optimizer slots
    n_train = sum(e['count'] for e in m._variable_layout() if e['trainable'])
    opt = Adam(lr=1e-4)
    opt.set_slots(3, np.ones(n_train, np.float32), np.ones(n_train, np.float32))
    c = ck.Checkpoint(m, opt, step=3)
    p2 = c.write(str(tmp_path / 'ckpt-3'))
    raw2 = T.read_bundle(p2, strings=True)
    g2 = pb.TrackableObjectGraph()
    g2.ParseFromString(raw2[T.OBJECT_GRAPH_KEY])
    opt_node = g2.nodes[follow(g2, 'optimizer')]
    kid = follow(g2, 'net/DownLayers/0/ConvLSTM/0/cell/kernel')
    slots = {(s.original_variable_node_id, s.slot_name): s.slot_variable_node_id for s in opt_node.slot_variables}
    assert (kid, 'm') in slots and (kid, 'v') in slots
    assert g2.nodes[slots[(kid, 'm')]].attributes[0].checkpoint_key == \
        'net/DownLayers/0/ConvLSTM/0/cell/kernel/.OPTIMIZER_SLOT/optimizer/m' + T.VAR_SUFFIX
    assert follow(g2, 'step') > 0 and follow(g2, 'optimizer/iter') > 0
    # a flipped byte inside the string tensor is caught by its checksum
    data = bytearray(open(p2 + '.data-00000-of-00001', 'rb').read())
    e = T._parse_entry(T.read_table(p2 + '.index')[T.OBJECT_GRAPH_KEY.encode()])
    data[e['offset'] + e['size'] - 3] ^= 0x40
    open(p2 + '.data-00000-of-00001', 'wb').write(bytes(data))
    with pytest.raises(ValueError):
        T.read_bundle(p2, strings=True)
