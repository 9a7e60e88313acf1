"""CPU checks of the boundary: the C-ABI library exports every symbol the header declares, the reference-facing Python
mirrors keep the reference's names / attributes, and the product refuses to run without a GPU (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from lstm_unet_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'lstm_unet_b200.h')).read()
    declared = set(re.findall(r'\b(lu_[a-z_0-9]+)\s*\(', header))
    assert declared and declared == set(_lib.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(_lib.default_library_path())
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.lu_is_cuda_build() == 1


def test_plan_and_param_layout_without_gpu():
    # lu_create touches no device: layer plan, Keras variable layout and workspace size are host logic
    from lstm_unet_b200 import _lib
    from oracle import lstm_unet_oracle as O
    lib = _lib.load_library()
    cfg = _lib.make_config(O.CTC_NET_PARAMS, 'NCHW', True, batch=4, max_t=8, height=512, width=512)
    h = ctypes.c_void_p()
    assert lib.lu_create(ctypes.byref(cfg), ctypes.byref(h)) == 0, lib.lu_last_error()
    nt, ne, ntr = ctypes.c_int32(), ctypes.c_int64(), ctypes.c_int64()
    lib.lu_param_count(h, ctypes.byref(nt), ctypes.byref(ne), ctypes.byref(ntr))
    assert (ne.value, ntr.value) == (74613059, 74606531)
    names = []
    buf = ctypes.create_string_buffer(256)
    for i in range(nt.value):
        lib.lu_param_info(h, i, buf, 256, None, None, None, None)
        names.append(buf.value.decode())
    assert names == [n for n, _, k in O.build_param_specs(O.CTC_NET_PARAMS) if k in O.TRAINABLE_KINDS] + \
        [n for n, _, k in O.build_param_specs(O.CTC_NET_PARAMS) if k not in O.TRAINABLE_KINDS]
    fl = ctypes.c_double()
    lib.lu_forward_flops(h, 1, ctypes.byref(fl))
    assert abs(fl.value / 3.3114e12 - 1) < 2e-3                     # 528x528 padded frame, SURVEY 8d
    nb = ctypes.c_size_t()
    lib.lu_workspace_bytes(h, ctypes.byref(nb))
    assert 1e9 < nb.value < 60e9
    lib.lu_destroy(h)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from lstm_unet_b200.Networks import ULSTMnet2D
    m = ULSTMnet2D()
    with pytest.raises(RuntimeError):
        m(np.zeros((1, 1, 1, 16, 16), np.float32), False)
    # the steps either side of the model have no CPU path either
    from lstm_unet_b200.postprocess import PostProcessor
    from lstm_unet_b200.augment import SequenceAugmenter
    from lstm_unet_b200 import losses
    from lstm_unet_b200.data import CTCRAMReaderSequence2D
    for make in (PostProcessor, SequenceAugmenter, lambda: losses.seg_measure(2)):
        with pytest.raises(RuntimeError):
            make()
    rd = CTCRAMReaderSequence2D(sequences=[{'images': np.zeros((4, 16, 16)), 'segs': np.zeros((4, 16, 16))}],
                                image_crop_size=(16, 16), unroll_len=2, batch_size=1)
    with pytest.raises(RuntimeError):
        rd.get_batch()


def test_reference_surface_names():
    from lstm_unet_b200 import Networks, Params, losses, train2D, Inference2D
    for n in ('DEFAULT_NET_DOWN_PARAMS', 'DownBlock2D', 'UpBlock2D', 'ULSTMnet2D'):
        assert hasattr(Networks, n)
    for n in ('reset_states_per_batch', 'get_states', 'set_states', 'trainable_variables', 'save_weights', 'load_weights'):
        assert hasattr(Networks.ULSTMnet2D, n)
    p = Params.CTCParams({'batch_size': 3, 'not_a_param': 1})
    assert p.batch_size == 3 and p.channel_axis == 1 and p.net_model is Networks.ULSTMnet2D
    assert p.net_kernel_params['lstm_kernels'][3] == [(5, 512)]
    image, seg, _, is_last = p.train_data_provider.get_batch()
    assert image.shape == (3, 4, 1, 128, 128) and seg.shape == image.shape and is_last.shape == (3,)
    assert set(np.unique(seg)) <= {-1.0, 0.0, 1.0, 2.0}
    assert Params.CTCInferenceParams({'data_format': 'NHWC'}).channel_axis == 3
    assert callable(train2D.train) and callable(Inference2D.inference) and callable(losses.WeightedCELoss)
    with pytest.raises(ValueError):
        Networks.ULSTMnet2D({'down_conv_kernels': [[(3, 4)]], 'lstm_kernels': [], 'up_conv_kernels': [[(1, 3)]]})


def test_inference_reader_mirror(tmp_path):
    """data.CTCInferenceReader (DataHandeling.py:1572-1596): sorted files, reversed warm-up prefix, per-frame z-score"""
    import cv2
    from lstm_unet_b200.data import CTCInferenceReader
    rng = np.random.default_rng(0)
    raw = [rng.integers(0, 4000, size=(12, 10)).astype(np.uint16) for _ in range(4)]
    for i, a in enumerate(raw):
        cv2.imwrite(str(tmp_path / ('t%03d.tif' % i)), a)
    frames = list(CTCInferenceReader(str(tmp_path), 't*.tif', pre_sequence_frames=2).dataset)
    order = [1, 0, 0, 1, 2, 3]
    assert len(frames) == 6
    for f, k in zip(frames, order):
        a = raw[k].astype(np.float32)
        want = (a - a.mean()) / a.std()
        assert f.dtype == np.float32 and np.allclose(f, want, rtol=1e-6, atol=1e-6)
    with pytest.raises(ValueError):
        CTCInferenceReader(str(tmp_path), 'nothing*.tif')


def test_training_reader_loads_ctc_folder(tmp_path):
    """data.CTCRAMReaderSequence2D._read_sequence_to_ram_ (DataHandeling.py:60-133) on a folder in the layout
    create_sequence_metadata.py writes: filelist rows (raw, seg or None, tra, fully-annotated flag), per-frame z-score,
    partially annotated frames get background = -1, frames without a segmentation file are all -1"""
    import pickle
    import cv2
    from lstm_unet_b200.data import CTCRAMReaderSequence2D
    rng = np.random.default_rng(0)
    os.makedirs(tmp_path / '01')
    os.makedirs(tmp_path / '01_GT' / 'SEG')
    rows, raws, segs = [], [], []
    for t in range(4):
        raw = rng.integers(0, 4000, size=(20, 24)).astype(np.uint16)
        cv2.imwrite(str(tmp_path / '01' / ('t%03d.tif' % t)), raw)
        raws.append(raw)
        if t == 3:
            rows.append((os.path.join('01', 't%03d.tif' % t), None, None, None))
            segs.append(None)
            continue
        seg = np.zeros((20, 24), np.uint16)
        seg[3:8, 4:9] = 1
        seg[12:17, 10:20] = 2
        cv2.imwrite(str(tmp_path / '01_GT' / 'SEG' / ('man_seg%03d.tif' % t)), seg)
        rows.append((os.path.join('01', 't%03d.tif' % t), os.path.join('01_GT', 'SEG', 'man_seg%03d.tif' % t), None, t != 1))
        segs.append(seg)
    with open(tmp_path / 'metadata_01.pickle', 'wb') as f:
        pickle.dump({'filelist': rows, 'max': 4000, 'min': 0, 'shape': (20, 24)}, f)
    rd = CTCRAMReaderSequence2D(sequence_folder_list=[(str(tmp_path), '01')], image_crop_size=(16, 16), unroll_len=2, batch_size=1,
                                seed=0)
    rd._read_sequence_to_ram_()
    seq = rd.sequence_data[(str(tmp_path), '01')]
    assert seq['images'].shape == (4, 20, 24) and list(seq['full_seg']) == [1, 0, 1, 0]
    for t in range(4):
        a = raws[t].astype(np.float32)
        assert np.allclose(seq['images'][t], (a - a.mean()) / a.std(), rtol=1e-5, atol=1e-5)
    assert np.array_equal(seq['segs'][0], segs[0])                       # fully annotated: labels as stored
    assert np.array_equal(seq['segs'][1], np.where(segs[1] == 0, -1.0, segs[1].astype(np.float64)))   # partially annotated: background unknown
    assert np.all(seq['segs'][3] == -1)                                  # no segmentation file


@pytest.mark.skipif(not os.path.exists('/root/reference/Params.py'), reason='the reference is only present in the build container')
def test_config_surface_equals_the_references():
    """The reference's Params.py, imported in a subprocess with stand-in tensorflow / matplotlib modules (class bodies only,
    nothing is instantiated): every configuration attribute of CTCParams / CTCInferenceParams exists in the mirror with the
    same default, except the documented differences (dry_run and save_intermediate default to off here)."""
    import subprocess
    import sys
    code = r'''
import sys, types
class _Any(types.ModuleType):
    def __init__(self, *a, **k): types.ModuleType.__init__(self, 'stub')
    def __getattr__(self, k):
        if k.startswith('__'): raise AttributeError(k)
        return _Any()
    def __call__(self, *a, **k): return _Any()
    def __mro_entries__(self, bases): return (object,)
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def __iter__(self): return iter(())
for m in ('tensorflow', 'tensorflow.python', 'tensorflow.python.keras', 'tensorflow.keras', 'matplotlib', 'matplotlib.pyplot'):
    sys.modules[m] = _Any()
sys.modules['tensorflow'].__version__ = '2.0'
sys.path.insert(0, '/root/reference')
import Params as RP
sys.path.insert(0, %r)
from lstm_unet_b200 import Params as MP
def attrs(c): return {k: v for k, v in vars(c).items() if not k.startswith('_') and not callable(v)}
bad = []
for name, allowed in (('CTCParams', {'dry_run'}), ('CTCInferenceParams', {'dry_run', 'save_intermediate'})):
    r, m = attrs(getattr(RP, name)), attrs(getattr(MP, name))
    bad += ['%%s.%%s missing' %% (name, k) for k in set(r) - set(m)]
    bad += ['%%s.%%s = %%r, reference %%r' %% (name, k, m[k], r[k]) for k in set(r) & set(m) if r[k] != m[k] and k not in allowed]
    assert set(m) - set(r) <= {'precision', 'seed', 'train_data_provider', 'val_data_provider'}, set(m) - set(r)
print('BAD:' + ';'.join(bad))
''' % ROOT
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == 'BAD:', r.stdout


_REF_STRUCTURE_CODE = r'''
import sys, types, json
class Rec:
    def __init__(self, kind, kw): self.kind, self.kw = kind, kw
class Layers:
    def __getattr__(self, kind):
        return lambda *a, **kw: Rec(kind, dict(kw, _args=list(a)))
class _Any(types.ModuleType):
    def __init__(self, *a, **k): types.ModuleType.__init__(self, 'stub')
    def __getattr__(self, k):
        if k.startswith('__'): raise AttributeError(k)
        return _Any()
    def __call__(self, *a, **k): return _Any()
    def __mro_entries__(self, bases): return (object,)
keras = _Any(); keras.layers = Layers()
tf = _Any(); tf.__version__ = '2.0'
sys.modules['tensorflow'] = tf
sys.modules['tensorflow.python'] = _Any()
sys.modules['tensorflow.python.keras'] = keras
tf.python = sys.modules['tensorflow.python']; tf.python.keras = keras
sys.path.insert(0, '/root/reference')
import Networks as RN
def describe(net, fmt):
    m = RN.ULSTMnet2D(net, fmt, True)
    d = {'total_stride': m.total_stride, 'last_depth': m.last_depth, 'softmax_axis': m.Softmax.kw['_args'][0], 'down': [], 'up': []}
    for b in m.DownLayers:
        d['down'].append({'total_stride': b.total_stride,
                          'lstm': [[l.kw['filters'], l.kw['kernel_size'], l.kw['strides'], l.kw['padding'], l.kw['return_sequences'], l.kw['stateful'], l.kw['data_format']] for l in b.ConvLSTM],
                          'conv': [[l.kw['filters'], l.kw['kernel_size'], l.kw['strides'], l.kw['padding'], l.kw['use_bias']] for l in b.Conv],
                          'bn_axis': [l.kw['axis'] for l in b.BN]})
    for b in m.UpLayers:
        d['up'].append({'up_factor': b.up_factor, 'return_logits': b.return_logits,
                        'conv': [[l.kw['filters'], l.kw['kernel_size'], l.kw['strides'], l.kw['padding'], l.kw['use_bias']] for l in b.Conv]})
    return d
nets = json.loads(sys.argv[1])
out = {'default': describe(RN.DEFAULT_NET_DOWN_PARAMS, 'NCHW'), 'default_params': RN.DEFAULT_NET_DOWN_PARAMS}
for name, (net, fmt) in nets.items():
    out[name] = describe(net, fmt)
try:
    RN.ULSTMnet2D({'down_conv_kernels': [[(3, 4)]], 'lstm_kernels': [], 'up_conv_kernels': [[(3, 4)]]})
    out['mismatch_raises'] = False
except ValueError:
    out['mismatch_raises'] = True
print('JSON:' + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.exists('/root/reference/Networks.py'), reason='the reference is only present in the build container')
def test_network_structure_equals_the_references_constructor():
    """The reference's ULSTMnet2D.__init__ (Networks.py:179-206) and block constructors (:37-58,124-139) executed in a
    subprocess against a RECORDING stand-in for Keras: the layers they construct (filters, kernel sizes, strides, padding,
    stateful / return_sequences flags, BN axis), the stride / up-factor / return_logits pattern, total_stride, last_depth
    and the soft-max axis -- against the mirror's block descriptors and the library's Keras-layout parameter shapes."""
    import json
    import subprocess
    import sys
    from lstm_unet_b200 import Networks, _lib
    from oracle import lstm_unet_oracle as O
    nets = {'ctc': (O.CTC_NET_PARAMS, 'NCHW'),
            'odd': ({'down_conv_kernels': [[(3, 6)], [(3, 10), (5, 10)]], 'lstm_kernels': [[(3, 5), (3, 7)], [(5, 9)]],
                     'up_conv_kernels': [[(3, 6)], [(3, 5), (1, 3)]]}, 'NHWC')}
    r = subprocess.run([sys.executable, '-c', _REF_STRUCTURE_CODE, json.dumps(nets)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    ref = json.loads([l for l in r.stdout.splitlines() if l.startswith('JSON:')][-1][5:])
    assert ref['mismatch_raises']
    with pytest.raises(ValueError):
        Networks.ULSTMnet2D({'down_conv_kernels': [[(3, 4)]], 'lstm_kernels': [], 'up_conv_kernels': [[(3, 4)]]})
    as_lists = lambda d: {k: [[list(t) for t in lvl] for lvl in v] for k, v in d.items()}
    assert as_lists(Networks.DEFAULT_NET_DOWN_PARAMS) == as_lists(ref['default_params'])
    lib = _lib.load_library()
    for name, (net, fmt) in list(nets.items()) + [('default', (Networks.DEFAULT_NET_DOWN_PARAMS, 'NCHW'))]:
        want = ref[name]
        m = Networks.ULSTMnet2D(net, fmt, True)
        L = len(net['down_conv_kernels'])
        assert m.last_depth == want['last_depth']
        total = 1
        for i, (b, wb) in enumerate(zip(m.DownLayers, want['down'])):
            assert b.total_stride == wb['total_stride'] == (2 if i < L - 1 else 1)
            total *= b.total_stride
            assert [[f, k] for k, f in b.lstm_kernels] == [[l[0], l[1]] for l in wb['lstm']]
            assert all(l[2:6] == [1, 'same', True, True] for l in wb['lstm'])                 # what the kernels implement
            assert [[f, k] for k, f in b.conv_kernels] == [[c[0], c[1]] for c in wb['conv']]
            assert [c[2] for c in wb['conv']] == [b.stride] + [1] * (len(wb['conv']) - 1)
            assert all(c[3] == 'same' and c[4] is True for c in wb['conv'])
            assert all(a == (1 if fmt[1] == 'C' else -1) for a in wb['bn_axis'])
        assert total == want['total_stride']
        for i, (b, wb) in enumerate(zip(m.UpLayers, want['up'])):
            assert b.up_factor == wb['up_factor'] == (2 if i > 0 else 1)
            assert bool(b.return_logits) == bool(wb['return_logits']) == (i == L - 1)
            assert [[f, k] for k, f in b.kernels] == [[c[0], c[1]] for c in wb['conv']]
            assert all(c[2] == 1 and c[3] == 'same' and c[4] is True for c in wb['conv'])
        assert want['softmax_axis'] == (2 if fmt[1] == 'C' else 0)                           # channel_axis + 1 (the NHWC quirk)
        # the library's Keras-layout parameter shapes carry the same filters / kernel sizes
        cfg = _lib.make_config(net, fmt, True, batch=1, max_t=1, height=32, width=32)
        h = ctypes.c_void_p()
        assert lib.lu_create(ctypes.byref(cfg), ctypes.byref(h)) == 0
        nt = ctypes.c_int32()
        lib.lu_param_count(h, ctypes.byref(nt), None, None)
        shapes, buf = {}, ctypes.create_string_buffer(256)
        for i in range(nt.value):
            shp, rank = (ctypes.c_int64 * 4)(), ctypes.c_int32()
            lib.lu_param_info(h, i, buf, 256, shp, ctypes.byref(rank), None, None)
            shapes[buf.value.decode()] = tuple(shp[j] for j in range(rank.value))
        lib.lu_destroy(h)
        for li, wb in enumerate(want['down']):
            for j, l in enumerate(wb['lstm']):
                ks = shapes['DownLayers/%d/ConvLSTM/%d/kernel' % (li, j)]
                assert ks[0] == ks[1] == l[1] and ks[3] == 4 * l[0]
                assert shapes['DownLayers/%d/ConvLSTM/%d/recurrent_kernel' % (li, j)] == (l[1], l[1], l[0], 4 * l[0])
            for j, c in enumerate(wb['conv']):
                ks = shapes['DownLayers/%d/Conv/%d/kernel' % (li, j)]
                assert ks[0] == ks[1] == c[1] and ks[3] == c[0]
        for ui, wb in enumerate(want['up']):
            for j, c in enumerate(wb['conv']):
                ks = shapes['UpLayers/%d/Conv/%d/kernel' % (ui, j)]
                assert ks[0] == ks[1] == c[1] and ks[3] == c[0]
                is_logits = wb['return_logits'] and j == len(wb['conv']) - 1
                assert (('UpLayers/%d/BN/%d/gamma' % (ui, j)) in shapes) == (not is_logits)     # Networks.py:148-149
