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
