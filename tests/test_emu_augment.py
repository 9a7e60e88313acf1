"""CPU tests of the device augmentation (lu_augment_sequence / lu_elastic_coords) through the TEST-ONLY host build:
against the vectors made by the reference's own helpers and against the oracle on seeded inputs.  Segmentations are
bit-exact; images agree to float32 rounding of the frame mean (the one value summed in a different order)."""
import numpy as np
import pytest

from oracle import augment_oracle as A
from tests.emu_backend import NumpyBackend, build_emu
from tests.test_augment_oracle import augment_cases


def emu_augmenter():
    from lstm_unet_b200 import _lib
    from lstm_unet_b200.augment import SequenceAugmenter
    lib = _lib.load_library(build_emu())
    assert lib.lu_is_cuda_build() == 0
    return SequenceAugmenter(_lib_override=lib, _backend=NumpyBackend())


@pytest.mark.parametrize('case', list(augment_cases()), ids=lambda c: c[0])
def test_emu_matches_reference_vectors(case):
    name, c = case
    aug = emu_augmenter()
    flip, rot = (int(c['flip_rot'][0]), int(c['flip_rot'][1])), int(c['flip_rot'][2])
    coords = c['coords'].reshape(-1).copy() if 'coords' in c else None
    img, seg = aug.augment(c['img'], c['seg'], c['contrast'], c['brightness'], c.get('affine'), coords, flip, rot)
    assert img.shape == c['out_img'].shape
    assert np.array_equal(seg, c['out_seg'])
    np.testing.assert_allclose(img, c['out_img'], rtol=2e-6, atol=2e-4)


@pytest.mark.parametrize('case', [c for c in augment_cases() if 'rand2' in c[1]], ids=lambda c: c[0])
def test_emu_elastic_field_matches_reference_vectors(case):
    name, c = case
    H, W = c['img'].shape[1:]
    coords = emu_augmenter().elastic_coords(c['rand2'], W * 2, W * 0.15).reshape(2, H, W)
    np.testing.assert_allclose(coords, c['coords'], rtol=0, atol=1e-10)


def test_emu_not_randomized_is_exact_and_errors():
    aug = emu_augmenter()
    imgs, segs = A.synthetic_sequence(2, 24, 30, 9)
    img, seg = aug.augment(imgs, segs, randomize=False, flip=(1, 1))
    for t in range(2):
        ri, rs = A.augment_frame(imgs[t], segs[t], 1, 0, None, None, (1, 1), 0, randomize=False)
        assert np.array_equal(img[t], ri) and np.array_equal(seg[t], rs)
    with pytest.raises(ValueError):
        aug.augment(imgs, segs, randomize=False, rot90=1)          # odd rotation of a non-square crop


def test_reader_mirror_batches_match_oracle_chain():
    """data.CTCRAMReaderSequence2D: the draws of a sequence (recorded in ``last_draws``) pushed through the oracle give
    the frames the reader hands out; batch layout, is_last semantics and the unroll windows follow DataHandeling.py:452-492"""
    from lstm_unet_b200.data import CTCRAMReaderSequence2D
    seqs = []
    for s in range(2):
        imgs, segs = A.synthetic_sequence(9, 40, 44, 20 + s, unlabeled_every=4)
        seqs.append({'images': imgs, 'segs': segs, 'full_seg': np.ones(9)})
    rd = CTCRAMReaderSequence2D(sequences=seqs, image_crop_size=(32, 32), unroll_len=2, batch_size=1, seed=3,
                                elastic_seed=5, _augmenter=emu_augmenter())
    rd.start_queues()
    image, seg, full, is_last = rd.get_batch()
    assert image.shape == (1, 2, 1, 32, 32) and seg.shape == (1, 2, 1, 32, 32) and full.shape == (1, 2) and is_last.shape == (1,)
    d = rd.last_draws
    src = seqs[d['key']]
    coords = A.elastic_coords(d['rand2'], 32 * 2, 32 * 0.15)
    n = len(d['idx'])
    assert n % 2 == 0 and n >= 2
    frames = [(image, seg, is_last)]
    for _ in range(n // 2 - 1):
        i2, s2, _, l2 = rd.get_batch()
        frames.append((i2, s2, l2))
    for w, (im, sg, last) in enumerate(frames):
        assert last[0] == (0.0 if w == n // 2 - 1 else 1.0)          # 0 = the window that ends the sequence
        for k in range(2):
            t = 2 * w + k
            f = d['idx'][t]
            ys, xs = slice(d['crop_y'], d['crop_y'] + 32), slice(d['crop_x'], d['crop_x'] + 32)
            ri, rs = A.augment_frame(src['images'][f][ys, xs], src['segs'][f][ys, xs], d['contrast'][t], d['brightness'][t],
                                     d['affine'], coords, d['flip'], d['rotate'])
            assert np.array_equal(sg[0, k, 0], rs)
            np.testing.assert_allclose(im[0, k, 0], ri, rtol=2e-6, atol=2e-4)
    # channels-last layout and the non-randomized path
    rd2 = CTCRAMReaderSequence2D(sequences=seqs, image_crop_size=(40, 44), unroll_len=3, batch_size=2, seed=1,
                                 data_format='NHWC', randomize=False, elastic_augmentation=False, _augmenter=emu_augmenter())
    image, seg, full, is_last = rd2.get_batch()
    assert image.shape == (2, 3, 40, 44, 1)
    assert np.array_equal(image[1, 0, :, :, 0], seqs[rd2.last_draws['key']]['images'][0])      # slot 1 was produced last
    assert set(np.unique(seg)) <= {-1.0, 0.0, 1.0, 2.0}


def test_emu_randomized_sweep_against_oracle():
    """seeded sweep over crop shapes, affine strengths, flips / rotations, with and without the elastic warp"""
    from lstm_unet_b200.augment import random_affine
    rng = np.random.RandomState(31)
    aug = emu_augmenter()
    for trial in range(10):
        square = trial % 2 == 0
        H = int(rng.randint(8, 40)); W = H if square else int(rng.randint(8, 40))
        T = int(rng.randint(1, 4))
        imgs, segs = A.synthetic_sequence(T, H, W, 300 + trial, unlabeled_every=2 if trial % 3 == 0 else 0)
        elastic = trial % 4 != 3
        affine = coords_dev = coords = None
        if elastic:
            affine = random_affine((H, W), W * rng.uniform(0.02, 0.2), rng)
            rand2 = np.stack([rng.rand(H, W), rng.rand(H, W)])
            coords = A.elastic_coords(rand2, W * 2, W * 0.15)
            coords_dev = aug.elastic_coords(rand2, W * 2, W * 0.15)
            np.testing.assert_allclose(coords_dev.reshape(2, H, W), coords, rtol=0, atol=1e-10)
        contrast = (rng.rand(T) + 0.5).astype(np.float32)
        brightness = ((rng.rand(T) - 0.5) * 0.2 * imgs.max()).astype(np.float32)
        flip, rot = (int(rng.randint(0, 2)), int(rng.randint(0, 2))), int(rng.randint(0, 4)) if square else 2 * int(rng.randint(0, 2))
        img, seg = aug.augment(imgs, segs, contrast, brightness, affine, coords_dev, flip, rot)
        for t in range(T):
            ri, rs = A.augment_frame(imgs[t], segs[t], contrast[t], brightness[t], affine, coords, flip, rot)
            assert np.array_equal(seg[t], rs), (trial, t)
            np.testing.assert_allclose(img[t], ri, rtol=2e-6, atol=2e-4)


REF_META = '/root/reference/metadata_files.tar.gz'


@pytest.mark.skipif(not __import__('os').path.exists(REF_META), reason='the reference archive is only present in the build container')
def test_reader_on_a_real_ctc_metadata_file(tmp_path):
    """The reference ships the metadata pickles of seven CTC datasets (metadata_files.tar.gz).  One of them, with frames
    synthesised at the paths it lists, goes through the reader mirror: folder loading (partial / missing annotations as in
    the real file list), the reference-order draws, device augmentation (host build here) and batching."""
    import os
    import pickle
    import tarfile
    import cv2
    from lstm_unet_b200.data import CTCRAMReaderSequence2D
    with tarfile.open(REF_META) as tf:
        member = next(m for m in tf.getmembers() if m.name.endswith('DIC-C2DH-HeLa/metadata_01.pickle'))
        meta = pickle.load(tf.extractfile(member))
    root = tmp_path / 'DIC-C2DH-HeLa'
    H, W = meta['shape']
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:H, 0:W]
    for t, (raw_name, seg_name, _, _flag) in enumerate(meta['filelist']):
        os.makedirs(os.path.dirname(root / raw_name), exist_ok=True)
        img = (rng.integers(0, 60, size=(H, W)) + 100 * (((yy - 200 - t) ** 2 + (xx - 250) ** 2) < 900)).astype(np.uint8)
        cv2.imwrite(str(root / raw_name), img)
        if seg_name is not None:
            os.makedirs(os.path.dirname(root / seg_name), exist_ok=True)
            seg = np.zeros((H, W), np.uint16)
            seg[((yy - 200 - t) ** 2 + (xx - 250) ** 2) < 900] = 1
            seg[((yy - 400) ** 2 + (xx - 100 - t) ** 2) < 400] = 2
            cv2.imwrite(str(root / seg_name), seg)
    with open(root / 'metadata_01.pickle', 'wb') as f:
        pickle.dump(meta, f)
    rd = CTCRAMReaderSequence2D(sequence_folder_list=[(str(root), '01')], image_crop_size=(64, 64), unroll_len=4, batch_size=2,
                                seed=1, elastic_seed=2, _augmenter=emu_augmenter())
    rd.start_queues()
    seq = rd.sequence_data[(str(root), '01')]
    n_seg = sum(r[1] is not None for r in meta['filelist'])
    assert seq['images'].shape == (len(meta['filelist']), H, W)
    assert int((seq['segs'].reshape(len(meta['filelist']), -1).max(1) > 0).sum()) == n_seg
    assert int(seq['full_seg'].sum()) == sum(bool(r[3]) for r in meta['filelist'])
    for _ in range(3):
        image, seg, full, is_last = rd.get_batch()
        assert image.shape == (2, 4, 1, 64, 64) and seg.shape == (2, 4, 1, 64, 64) and full.shape == (2, 4) and is_last.shape == (2,)
        assert np.isfinite(image).all() and set(np.unique(seg)) <= {-1.0, 0.0, 1.0, 2.0}
    # the reference's own loader (DataHandeling.py:60-133, imported with stand-in tensorflow / utils modules) on the same
    # folder gives the same arrays
    from tests.golden.make_augment_golden import reference_reader
    R = reference_reader()
    ref = R(sequence_folder_list=[(str(root), '01')], image_crop_size=(64, 64), unroll_len=4, batch_size=2, num_threads=1)
    ref._read_sequence_to_ram_()
    ref_seq = ref.sequence_data[(str(root), '01')]
    assert np.array_equal(ref_seq['images'], seq['images'])
    assert np.array_equal(ref_seq['segs'], seq['segs'])
    assert np.array_equal(ref_seq['full_seg'], seq['full_seg'])


@pytest.mark.skipif(not __import__('os').path.exists('/root/reference/DataHandeling.py'),
                    reason='the reference is only present in the build container')
def test_reader_sequence_equals_the_references_own_loop(tmp_path):
    """One whole pass of the reference's ``_load_and_enqueue`` (DataHandeling.py:262-428: sequence choice, sub-sampling,
    reversal, crop, flips, rotation, per-frame contrast / brightness draws, segmentation relabelling, queue payload) --
    executed from /root/reference with stand-in tensorflow queues and seeded ``random`` / ``np.random`` -- against the
    reader mirror seeded the same way.  Elastic augmentation is off here: the reference seeds that RandomState from OS
    entropy (:159), its arithmetic is pinned separately (tests/golden/augment.npz)."""
    import os
    import pickle
    import random
    import types
    import cv2
    from lstm_unet_b200.data import CTCRAMReaderSequence2D
    from tests.golden.make_augment_golden import reference_reader
    rng = np.random.default_rng(3)
    root = tmp_path / 'seq'
    os.makedirs(root / '01')
    os.makedirs(root / '01_GT' / 'SEG')
    rows = []
    yy, xx = np.mgrid[0:40, 0:48]
    for t in range(11):
        raw = (rng.integers(0, 50, size=(40, 48)) + 150 * (((yy - 12 - t) ** 2 + (xx - 20) ** 2) < 40)).astype(np.uint8)
        cv2.imwrite(str(root / '01' / ('t%03d.tif' % t)), raw)
        seg_name = None
        if t % 4 != 3:
            seg = np.zeros((40, 48), np.uint16)
            seg[((yy - 12 - t) ** 2 + (xx - 20) ** 2) < 40] = 1
            seg[((yy - 30) ** 2 + (xx - 35 + t) ** 2) < 30] = 2
            seg_name = os.path.join('01_GT', 'SEG', 'man_seg%03d.tif' % t)
            cv2.imwrite(str(root / seg_name), seg)
        rows.append((os.path.join('01', 't%03d.tif' % t), seg_name, None, (t % 2 == 0) if seg_name else None))
    with open(root / 'metadata_01.pickle', 'wb') as f:
        pickle.dump({'filelist': rows, 'shape': (40, 48), 'max': 255, 'min': 0}, f)
    folders = [(str(root), '01')]
    kw = dict(image_crop_size=(16, 16), unroll_len=3, deal_with_end=0, batch_size=1, data_format='NCHW', randomize=True,
              elastic_augmentation=False)
    for seed in (0, 1, 2, 5):
        # ---- the reference's loop, one sequence ----
        R = reference_reader()
        ref = R(sequence_folder_list=folders, num_threads=3, **kw)
        random.seed(seed); np.random.seed(seed)
        ref._read_sequence_to_ram_()
        got = []
        ref.coord = types.SimpleNamespace(should_stop=lambda: len(got) > 0)      # stop once one sequence is enqueued
        q = types.SimpleNamespace(enqueue_many=lambda payload: got.append(payload))
        q_stat = lambda: types.SimpleNamespace(numpy=lambda: 0.0)
        ref._load_and_enqueue(q, q_stat)
        assert len(got) == 1
        r_img, r_seg, r_full, r_last, _names = got[0]
        # ---- the mirror, same seeds, same call order ----
        random.seed(seed); np.random.seed(seed)
        rd = CTCRAMReaderSequence2D(sequence_folder_list=folders, _augmenter=emu_augmenter(), **kw)
        rd.start_queues()
        rd._produce(0)
        m_img, m_seg, m_full, m_last, _ = rd._fifo[0][0]
        assert len(r_img) == m_img.shape[0] and len(r_img) % 3 == 0
        assert np.array_equal(np.stack(r_seg), m_seg), seed
        np.testing.assert_allclose(m_img, np.stack(r_img), rtol=2e-6, atol=2e-6)
        assert np.array_equal(np.asarray(r_full, np.float32), m_full) and np.array_equal(np.asarray(r_last, np.float32), m_last)
