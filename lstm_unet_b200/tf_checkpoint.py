"""TensorFlow-free reader / writer of TF2 checkpoints ("tensor bundle": ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``)
for the weights of ``ULSTMnet2D`` -- SURVEY 8f row 1: the on-disk format either side of the hot path
(reference: ``model.save_weights(model_fname, save_format='tf')`` train2D.py:236, ``model.load_weights(...)``
Inference2D.py:34, ``tf.train.Checkpoint(step, optimizer, net=model)`` train2D.py:62).

Formats implemented from their published specifications (TensorFlow is not installable here, so this module is validated by
round trips, by the CRC-32C and Snappy known-answer vectors, and structurally -- not yet against a file written by
TensorFlow itself; say so when reporting parity):

* ``.index``: an SSTable in the LevelDB table format (data blocks of prefix-compressed key/value entries with restart
  arrays, block trailers = 1 compression byte + masked CRC-32C, index block, 48-byte footer with magic
  0xdb4775248b80fb57); blocks may be Snappy-compressed.  Key "" -> ``BundleHeaderProto``; every other key ->
  ``BundleEntryProto`` {dtype=1, shape=2, shard_id=3, offset=4, size=5, crc32c=6}.
* ``.data-XXXXX-of-YYYYY``: raw little-endian tensor bytes at (offset, size).
* Keras object-graph variable keys: ``<attr path>/.ATTRIBUTES/VARIABLE_VALUE`` -- e.g.
  ``DownLayers/0/ConvLSTM/0/cell/kernel/.ATTRIBUTES/VARIABLE_VALUE`` -- optionally prefixed with ``net/`` when written
  through ``tf.train.Checkpoint(net=model)``.
* ``_CHECKPOINTABLE_OBJECT_GRAPH``: a scalar DT_STRING tensor holding a serialized ``TrackableObjectGraph`` proto
  (nodes with ``children`` edges {node_id, local_name}, ``attributes`` {name, full_name, checkpoint_key} and the
  optimizer's ``slot_variables``) -- what Keras' object-based ``load_weights`` / ``tf.train.Checkpoint.restore`` match
  against the live objects.  Written as the trie of the variable keys' attribute paths; string tensors are stored as
  [varint64 lengths][masked CRC-32C of the lengths as uint32][bytes] (tensor_bundle.cc, WriteStringTensor).
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_STRING, DT_INT64 = 1, 2, 3, 7, 9
OBJECT_GRAPH_KEY = '_CHECKPOINTABLE_OBJECT_GRAPH'
_NP_OF_DT = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_INT64: np.int64}
VAR_SUFFIX = '/.ATTRIBUTES/VARIABLE_VALUE'

# ---------------------------------------------------------------------------------------------------- CRC-32C (Castagnoli)
_CRC_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ 0x82F63B78 if _c & 1 else _c >> 1
    _CRC_TABLE.append(_c)
_CRC_NP = np.array(_CRC_TABLE, dtype=np.uint32)


def crc32c(data, crc=0):
    crc ^= 0xFFFFFFFF
    for b in bytes(data):
        crc = _CRC_TABLE[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def mask_crc(crc):
    """leveldb / TF crc masking: rotate right by 15 and add a constant."""
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------- varints / protobuf
def _put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _parse_proto(buf):
    """Minimal protobuf wire parser -> list of (field_number, wire_type, value)."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]; pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        out.append((f, wt, v))
    return out


def _parse_entry(buf):
    e = {'dtype': 0, 'shape': [], 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': None}
    for f, wt, v in _parse_proto(buf):
        if f == 1:
            e['dtype'] = v
        elif f == 2:                                   # TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}
            for f2, _, v2 in _parse_proto(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _parse_proto(v2):
                        if f3 == 1:
                            size = v3
                    e['shape'].append(size)
        elif f == 3:
            e['shard_id'] = v
        elif f == 4:
            e['offset'] = v
        elif f == 5:
            e['size'] = v
        elif f == 6:
            e['crc32c'] = struct.unpack('<I', v)[0]
    return e


def _encode_entry(dtype, shape, shard_id, offset, size, crc_masked):
    dims = b''.join(b'\x12' + _put_varint(len(d)) + d for d in (b'\x08' + _put_varint(s) for s in shape))
    out = b'\x08' + _put_varint(dtype)
    out += b'\x12' + _put_varint(len(dims)) + dims
    if shard_id:
        out += b'\x18' + _put_varint(shard_id)
    if offset:
        out += b'\x20' + _put_varint(offset)
    out += b'\x28' + _put_varint(size)
    out += b'\x35' + struct.pack('<I', crc_masked)
    return out


# ---------------------------------------------------------------------------------------------------- Snappy (decode only)
def snappy_uncompress(buf):
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]; pos += 1
        kind = tag & 3
        if kind == 0:                                  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], 'little'); pos += nb
            ln += 1
            out += buf[pos:pos + ln]; pos += ln
            continue
        if kind == 1:
            ln = 4 + ((tag >> 2) & 7)
            off = ((tag >> 5) << 8) | buf[pos]; pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = buf[pos] | (buf[pos + 1] << 8); pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], 'little'); pos += 4
        if off == 0 or off > len(out):
            raise ValueError('corrupt snappy stream')
        for _ in range(ln):                            # byte-wise: copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError('snappy length mismatch: %d != %d' % (len(out), n))
    return bytes(out)


# ---------------------------------------------------------------------------------------------------- SSTable
def _read_block(data, offset, size, verify=True):
    contents = data[offset:offset + size]
    ctype = data[offset + size]
    stored = struct.unpack('<I', data[offset + size + 1:offset + size + 5])[0]
    if verify and mask_crc(crc32c(data[offset:offset + size + 1])) != stored:
        raise ValueError('block checksum mismatch at offset %d' % offset)
    if ctype == 1:
        contents = snappy_uncompress(contents)
    elif ctype != 0:
        raise ValueError('unknown block compression %d' % ctype)
    return contents


def _block_entries(block):
    num_restarts = struct.unpack('<I', block[-4:])[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos, key, out = 0, b'', []
    while pos < limit:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]; pos += non_shared
        out.append((key, block[pos:pos + vlen])); pos += vlen
    return out


def read_table(path, verify=True):
    data = open(path, 'rb').read()
    if len(data) < 48 or struct.unpack('<Q', data[-8:])[0] != TABLE_MAGIC:
        raise ValueError('%s is not an SSTable (bad magic)' % path)
    footer = data[-48:]
    _, p = _get_varint(footer, 0)            # metaindex handle
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    entries = {}
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, q = _get_varint(handle, 0)
        bsize, _ = _get_varint(handle, q)
        for k, v in _block_entries(_read_block(data, boff, bsize, verify)):
            entries[k] = v
    return entries


def _build_block(items, restart_interval=16):
    out, restarts, prev = bytearray(), [], b''
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack('<I', r)
    out += struct.pack('<I', len(restarts))
    return bytes(out)


def write_table(path, items, block_size=4096):
    """items: sorted list of (key bytes, value bytes).  Uncompressed blocks (compression type 0)."""
    f = bytearray()

    def emit(block):
        off = len(f)
        f.extend(block + b'\x00')
        f.extend(struct.pack('<I', mask_crc(crc32c(block + b'\x00'))))
        return off, len(block)

    index, cur, cur_bytes = [], [], 0
    for k, v in items:
        cur.append((k, v)); cur_bytes += len(k) + len(v) + 6
        if cur_bytes >= block_size:
            off, size = emit(_build_block(cur))
            index.append((cur[-1][0], _put_varint(off) + _put_varint(size)))
            cur, cur_bytes = [], 0
    if cur:
        off, size = emit(_build_block(cur))
        index.append((cur[-1][0], _put_varint(off) + _put_varint(size)))
    moff, msize = emit(_build_block([]))
    ioff, isize = emit(_build_block(index, restart_interval=1))
    footer = _put_varint(moff) + _put_varint(msize) + _put_varint(ioff) + _put_varint(isize)
    footer += b'\x00' * (40 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    f.extend(footer)
    with open(path, 'wb') as fh:
        fh.write(bytes(f))


# ---------------------------------------------------------------------------------------------------- tensor bundle
def _string_tensor_bytes(strings):
    """On-disk form of a DT_STRING tensor and its (unmasked) running CRC: [varint64 len]*, masked CRC-32C of the lengths
    (each taken as a little-endian uint32, uint64 above 2^32), then the bytes (tensor_bundle.cc: WriteStringTensor)."""
    lengths, crc = b'', 0
    for st in strings:
        lengths += _put_varint(len(st))
        crc = crc32c(struct.pack('<Q' if len(st) > 0xFFFFFFFF else '<I', len(st)), crc)
    cks = struct.pack('<I', mask_crc(crc))
    crc = crc32c(cks, crc)
    for st in strings:
        crc = crc32c(st, crc)
    return lengths + cks + b''.join(strings), crc


def _decode_string_tensor(raw, entry, verify, key):
    n = 1
    for d in entry['shape']:
        n *= d
    pos, lens = 0, []
    for _ in range(n):
        ln, pos = _get_varint(raw, pos)
        lens.append(ln)
    pos += 4
    out = []
    for ln in lens:
        out.append(bytes(raw[pos:pos + ln])); pos += ln
    if verify and entry['crc32c'] is not None:
        _, crc = _string_tensor_bytes(out)
        if mask_crc(crc) != entry['crc32c']:
            raise ValueError('string tensor %s: checksum mismatch' % key)
    return out[0] if not entry['shape'] else out


def read_bundle(prefix, verify=True, strings=False):
    """-> {tensor key: numpy array} for the numeric tensors of a TF2 checkpoint; with ``strings`` also the DT_STRING
    tensors (bytes for a scalar, e.g. the serialized object graph under OBJECT_GRAPH_KEY)."""
    entries = read_table(prefix + '.index', verify)
    header = entries.pop(b'', None)
    num_shards = 1
    if header is not None:
        for f, _, v in _parse_proto(header):
            if f == 1:
                num_shards = v
            if f == 2 and v != 0:
                raise ValueError('big-endian tensor bundles are not supported')
    shards = {}
    out = {}
    for key, val in entries.items():
        e = _parse_entry(val)
        if e['dtype'] not in _NP_OF_DT and not (strings and e['dtype'] == DT_STRING):
            continue
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = open('%s.data-%05d-of-%05d' % (prefix, sid, num_shards), 'rb').read()
        raw = shards[sid][e['offset']:e['offset'] + e['size']]
        if e['dtype'] == DT_STRING:
            out[key.decode()] = _decode_string_tensor(raw, e, verify, key.decode())
            continue
        if verify and e['crc32c'] is not None and mask_crc(crc32c(raw)) != e['crc32c']:
            raise ValueError('tensor %s: checksum mismatch' % key.decode())
        out[key.decode()] = np.frombuffer(raw, dtype=_NP_OF_DT[e['dtype']]).reshape(e['shape']).copy()
    return out


def _pb_field(num, payload):
    return _put_varint((num << 3) | 2) + _put_varint(len(payload)) + payload


def encode_object_graph(var_keys):
    """Serialized ``TrackableObjectGraph`` for a set of variable checkpoint keys: node 0 is the root, every component of a
    key's attribute path is a ``children`` edge, the leaf carries the ``VARIABLE_VALUE`` attribute with its checkpoint key;
    ``<var>/.OPTIMIZER_SLOT/<optimizer path>/<slot>`` keys become slot-variable nodes referenced from the optimizer node
    (trackable_object_graph.proto: nodes = 1; children = 1 {node_id = 1, local_name = 2}; attributes = 2 {name = 1,
    full_name = 2, checkpoint_key = 3}; slot_variables = 3 {original_variable_node_id = 1, slot_name = 2,
    slot_variable_node_id = 3})."""
    nodes = [{'children': {}, 'attrs': [], 'slots': []}]

    def walk(path):
        cur = 0
        for part in path:
            nxt = nodes[cur]['children'].get(part)
            if nxt is None:
                nodes.append({'children': {}, 'attrs': [], 'slots': []})
                nxt = len(nodes) - 1
                nodes[cur]['children'][part] = nxt
            cur = nxt
        return cur
    slot_keys = []
    for key in sorted(var_keys):
        obj = key[:-len(VAR_SUFFIX)]
        if '/.OPTIMIZER_SLOT/' in obj:
            slot_keys.append(key)
            continue
        nid = walk(obj.split('/'))
        nodes[nid]['attrs'].append(('VARIABLE_VALUE', obj, key))
    for key in slot_keys:
        obj = key[:-len(VAR_SUFFIX)]
        var_path, rest = obj.split('/.OPTIMIZER_SLOT/')
        opt_path, slot_name = rest.rsplit('/', 1)
        var_id, opt_id = walk(var_path.split('/')), walk(opt_path.split('/'))
        nodes.append({'children': {}, 'attrs': [('VARIABLE_VALUE', var_path + '/' + slot_name, key)], 'slots': []})
        nodes[opt_id]['slots'].append((var_id, slot_name, len(nodes) - 1))
    out = b''
    for n in nodes:
        body = b''
        for name, nid in n['children'].items():
            body += _pb_field(1, b'\x08' + _put_varint(nid) + _pb_field(2, name.encode()))
        for name, full, ckey in n['attrs']:
            body += _pb_field(2, _pb_field(1, name.encode()) + _pb_field(2, full.encode()) + _pb_field(3, ckey.encode()))
        for var_id, slot_name, slot_id in n['slots']:
            body += _pb_field(3, b'\x08' + _put_varint(var_id) + _pb_field(2, slot_name.encode()) + b'\x18' + _put_varint(slot_id))
        out += _pb_field(1, body)
    return out


def write_bundle(prefix, tensors, object_graph=True):
    """tensors: {key: array} (float32, or int64 for integer arrays; bytes = a scalar string tensor).  One shard.  With
    ``object_graph`` the serialized ``TrackableObjectGraph`` of the variable keys is stored under OBJECT_GRAPH_KEY, which is
    what Keras' object-based ``load_weights`` / ``tf.train.Checkpoint.restore`` need besides the tensors; name-based readers
    (``tf.train.load_checkpoint(prefix).get_tensor(key)``) use the keys alone."""
    data, items, offset = bytearray(), [], 0
    header = b'\x08\x01' + b'\x1a\x02\x08\x01'        # num_shards = 1, version { producer: 1 }
    items.append((b'', header))
    tensors = dict(tensors)
    if object_graph and OBJECT_GRAPH_KEY not in tensors:
        tensors[OBJECT_GRAPH_KEY] = encode_object_graph([k for k in tensors if k.endswith(VAR_SUFFIX)])
    for key in sorted(tensors):
        if isinstance(tensors[key], (bytes, bytearray)):              # scalar DT_STRING (the object graph)
            raw, crc = _string_tensor_bytes([bytes(tensors[key])])
            items.append((key.encode(), _encode_entry(DT_STRING, (), 0, offset, len(raw), mask_crc(crc))))
            data += raw
            offset += len(raw)
            continue
        a = np.asarray(tensors[key])
        dt = DT_INT64 if a.dtype.kind in 'iu' else DT_FLOAT          # step counters are int64 in TF checkpoints
        a = np.ascontiguousarray(a, dtype=_NP_OF_DT[dt])
        raw = a.tobytes()
        items.append((key.encode(), _encode_entry(dt, a.shape, 0, offset, len(raw), mask_crc(crc32c(raw)))))
        data += raw
        offset += len(raw)
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    write_table(prefix + '.index', items)
    with open(prefix + '.data-00000-of-00001', 'wb') as fh:
        fh.write(bytes(data))


# ---------------------------------------------------------------------------------------------------- Keras <-> library names
def keras_key(name, prefix=''):
    """library / Keras variable name -> object-graph checkpoint key (ConvLSTM2D variables live in ``cell``)."""
    parts = name.split('/')
    if 'ConvLSTM' in parts:
        parts.insert(len(parts) - 1, 'cell')
    return prefix + '/'.join(parts) + VAR_SUFFIX


def load_model_weights(prefix, expected_names, verify=True):
    """Read the ULSTMnet2D variables from a TF2 checkpoint written by ``model.save_weights`` (keys at the root) or by
    ``tf.train.Checkpoint(net=model)`` (keys under ``net/``).  -> {library name: array}; KeyError lists what is missing."""
    tensors = read_bundle(prefix, verify)
    for root in ('', 'net/', 'model/'):
        if all(keras_key(n, root) in tensors for n in expected_names):
            return {n: tensors[keras_key(n, root)] for n in expected_names}
    missing = [keras_key(n) for n in expected_names if keras_key(n) not in tensors]
    raise KeyError('checkpoint %s lacks %d variables, e.g. %s' % (prefix, len(missing), missing[:3]))


def save_model_weights(prefix, named, root=''):
    write_bundle(prefix, {keras_key(n, root): v for n, v in named.items()})
