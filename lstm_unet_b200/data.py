"""Synthetic stand-in for the reference's data providers (DataHandeling.CTCRAMReaderSequence2D, out of scope):
same protocol -- ``start_queues(coord)`` and ``get_batch() -> (image, seg, full_seg, is_last)`` with
image (B,T,1,H,W) z-scored floats, seg (B,T,1,H,W) floats in {-1,0,1,2}, is_last (B,) 1 = sequence continues,
0 = last window -> reset (DataHandeling.py:378,471,528)."""
import numpy as np


class SyntheticSequenceProvider:
    def __init__(self, sequence_folder_list=None, image_crop_size=(128, 128), unroll_len=4, deal_with_end=0,
                 batch_size=2, queue_capacity=0, data_format='NCHW', randomize=True, return_dist=False,
                 num_threads=0, seed=0, sequence_len=5):
        self.crop, self.T, self.B = tuple(image_crop_size), unroll_len, batch_size
        self.channels_first = data_format[1] == 'C'
        self.rng = np.random.default_rng(seed)
        self.sequence_len = sequence_len
        self._count = 0

    def start_queues(self, coord=None, debug=False):
        return []

    def get_batch(self):
        H, W = self.crop
        shape = (self.B, self.T, 1, H, W) if self.channels_first else (self.B, self.T, H, W, 1)
        image = self.rng.standard_normal(shape).astype(np.float32)
        seg = self.rng.integers(-1, 3, size=shape).astype(np.float32)
        self._count += 1
        is_last = np.ones(self.B, dtype=np.float32)
        if self._count % self.sequence_len == 0:
            is_last[self.rng.integers(0, self.B)] = 0.0
        return image, seg, seg.copy(), is_last


class CTCRAMReaderSequence2D:
    """Mirror of the reference's training reader (DataHandeling.CTCRAMReaderSequence2D, DataHandeling.py:21-530): same
    constructor arguments and provider protocol, sequences held in RAM, random parameters drawn on the host in the
    reference's order (``_load_and_enqueue``, :262-330) -- and the per-frame augmentation chain (:330-380), which the
    reference runs frame by frame in Python worker threads, executed on the device for a whole sequence at a time
    (augment.SequenceAugmenter -> lu_augment_sequence).  ``get_batch()`` returns device tensors, so the batch never
    crosses PCIe after the sequence was uploaded once.

    Sequences come from ``sequence_folder_list`` (folders with ``metadata_<seq>.pickle`` and image files, read like
    ``_read_sequence_to_ram_`` :60-133; needs OpenCV) or, for tests and synthetic runs, from ``sequences``: a list of
    dicts ``{'images': (N,H,W), 'segs': (N,H,W), 'full_seg': (N,)}``.

    Differences from the reference, all on the host side: one producer (no TF queues / threads: a slot's next sequence
    is augmented when its FIFO runs short), and the elastic field's RandomState is seeded from ``elastic_seed`` when
    given (the reference seeds it from OS entropy, :159)."""

    def __init__(self, sequence_folder_list=None, image_crop_size=(128, 128), unroll_len=7, deal_with_end=0,
                 batch_size=4, queue_capacity=32, num_threads=3, data_format='NCHW', randomize=True, return_dist=False,
                 keep_sample=1, elastic_augmentation=True, sequences=None, seed=None, elastic_seed=None,
                 _augmenter=None):
        if return_dist:
            raise ValueError('return_dist (the distance-map target of another model family) is not built')
        self.sub_seq_size = tuple(image_crop_size)
        self.unroll_len, self.deal_with_end, self.batch_size = unroll_len, deal_with_end, batch_size
        self.data_format, self.randomize, self.keep_sample = data_format, randomize, keep_sample
        self.elastic_augmentation = elastic_augmentation
        self.sequence_folder_list = list(sequence_folder_list or [])
        self.sequence_data = {}
        if sequences is not None:
            for i, s in enumerate(sequences):
                self.sequence_data[i] = {'images': np.asarray(s['images']), 'segs': np.asarray(s['segs']),
                                         'full_seg': np.asarray(s.get('full_seg', np.ones(len(s['images']))))}
            self.sequence_folder_list = list(self.sequence_data)
        import random
        self._py_random = random.Random(seed) if seed is not None else random
        self._np_random = np.random.RandomState(seed) if seed is not None else np.random
        self._elastic_seed = elastic_seed
        self._aug = _augmenter
        self._fifo = [[] for _ in range(batch_size)]        # per batch slot: [img (n,H,W), seg, full_seg, is_last, cursor]
        self.last_draws = None

    # ---- DataHandeling.py:60-133 ---------------------------------------------------------------------------------
    def _read_sequence_to_ram_(self):
        import os
        import pickle
        for sequence_folder in self.sequence_folder_list:
            if sequence_folder in self.sequence_data:
                continue
            import cv2
            train_set, seq, folder = True, None, sequence_folder
            if isinstance(sequence_folder, tuple):
                if len(sequence_folder) == 2:
                    folder, seq = sequence_folder
                elif len(sequence_folder) == 3:
                    folder, seq, train_set = sequence_folder
            with open(os.path.join(folder, 'metadata_{}.pickle'.format(seq)), 'rb') as fobj:
                metadata = pickle.load(fobj)
            filename_list, img_size = metadata['filelist'], metadata['shape']
            if len(img_size) == 3:
                img_size = img_size[1:]
            n = len(filename_list)
            all_images, all_seg, all_full = np.zeros((n,) + tuple(img_size)), np.zeros((n,) + tuple(img_size)), np.zeros(n)
            for t, filename in enumerate(filename_list):
                img = cv2.imread(os.path.join(folder, filename[0]), -1)
                if img is None:
                    raise ValueError('Could not load image: {}'.format(os.path.join(folder, filename[0])))
                img = img.astype(np.float32)
                img = (img - img.mean()) / (img.std())
                full_seg = 1 if filename[3] is True else 0
                keep_seg = (self._np_random.rand() < self.keep_sample) and train_set
                full_seg = full_seg if keep_seg else 0
                if filename[1] is None or not keep_seg:
                    seg = np.ones(img.shape[:2]) * (-1)
                else:
                    seg = cv2.imread(os.path.join(folder, filename[1]), -1)
                    if not full_seg:
                        if seg is None:
                            seg, full_seg = np.ones(img.shape[:2]) * (-1), -1
                        else:
                            seg = seg.astype(np.float32)
                        seg[seg == 0] = -1
                all_images[t], all_seg[t], all_full[t] = img, seg, full_seg
            self.sequence_data[sequence_folder] = {'images': all_images, 'segs': all_seg, 'full_seg': all_full}

    def start_queues(self, coord=None, debug=False):
        self._read_sequence_to_ram_()
        if self._aug is None:
            from .augment import SequenceAugmenter
            self._aug = SequenceAugmenter()
        return []

    # ---- DataHandeling.py:262-330: the draws of one sequence, in the reference's order -----------------------------
    def _draw_sequence(self):
        from .augment import random_affine
        rs = self._np_random
        key = self._py_random.choice(self.sequence_folder_list)
        seq = self.sequence_data[key]
        n, img_size = len(seq['images']), seq['images'].shape[1:]
        sub = self.sub_seq_size
        d = {'key': key}
        d['sub_sample'] = rs.randint(1, 4) if self.randomize else 0
        d['reverse'] = rs.randint(0, 2) if self.randomize else 0
        d['crop_y'] = (rs.randint(0, img_size[0] - sub[0]) if self.randomize else 0) if img_size[0] - sub[0] > 0 else 0
        d['crop_x'] = (rs.randint(0, img_size[1] - sub[1]) if self.randomize else 0) if img_size[1] - sub[1] > 0 else 0
        d['flip'] = tuple(int(v) for v in (rs.randint(0, 2, 2) if self.randomize else [0, 0]))
        d['rotate'] = int(rs.randint(0, 4)) if self.randomize else 0
        if self.elastic_augmentation:
            state = np.random.RandomState(self._elastic_seed)
            d['affine'] = random_affine(sub, sub[1] * 0.08, state)
            d['rand2'] = np.stack([state.rand(*sub), state.rand(*sub)])     # x field first (:184-186)
        idx = list(range(n))
        if d['reverse']:
            idx.reverse()
        if d['sub_sample']:
            idx = idx[::d['sub_sample']]
        remainder = len(idx) % self.unroll_len
        if remainder:
            if self.deal_with_end == 0:
                idx = idx[:-remainder]
            elif self.deal_with_end == 1:
                idx += idx[-2:-self.unroll_len + remainder - 2:-1]
            elif self.deal_with_end == 2:
                idx += idx[-1:] * (self.unroll_len - remainder)
        d['idx'] = idx
        img_max = seq['images'].max()
        if self.randomize:
            cb = [(rs.rand() + 0.5, (rs.rand() - 0.5) * 0.2 * img_max) for _ in idx]
            d['contrast'] = np.array([c for c, _ in cb], np.float32)
            d['brightness'] = np.array([b for _, b in cb], np.float32)
        return d, seq

    def _produce(self, slot):
        d, seq = self._draw_sequence()
        sub = self.sub_seq_size
        tries = 0
        while not d['idx']:                                  # sequence shorter than one unroll window: draw another
            tries += 1
            if tries > 64:
                raise ValueError('no sequence is long enough for unroll_len=%d' % self.unroll_len)
            d, seq = self._draw_sequence()
        ys, xs = slice(d['crop_y'], d['crop_y'] + sub[0]), slice(d['crop_x'], d['crop_x'] + sub[1])
        img = np.ascontiguousarray(seq['images'][d['idx'], ys, xs], dtype=np.float32)
        seg = np.ascontiguousarray(seq['segs'][d['idx'], ys, xs], dtype=np.float32)
        coords = None
        if self.elastic_augmentation:
            coords = self._aug.elastic_coords(d['rand2'], sub[1] * 2, sub[1] * 0.15)
        o_img, o_seg = self._aug.augment(img, seg, d.get('contrast'), d.get('brightness'), d.get('affine'), coords,
                                         d['flip'], d['rotate'], randomize=self.randomize)
        n = len(d['idx'])
        full = np.maximum(0, np.asarray(seq['full_seg'])[d['idx']]).astype(np.float32)
        is_last = np.array([1.0 if (t + 1) < n else 0.0 for t in range(n)], np.float32)
        self._fifo[slot].append([o_img, o_seg, full, is_last, 0])
        self.last_draws = d

    def _dequeue_many(self, slot):
        T = self.unroll_len
        while not self._fifo[slot]:
            self._produce(slot)
        e = self._fifo[slot][0]
        c = e[4]
        out = (e[0][c:c + T], e[1][c:c + T], e[2][c:c + T], e[3][c:c + T])
        e[4] += T
        if e[4] >= len(e[2]):
            self._fifo[slot].pop(0)
        return out

    # ---- DataHandeling.py:452-492 ----------------------------------------------------------------------------------
    def get_batch(self):
        if self._aug is None:
            self.start_queues()
        parts = [self._dequeue_many(b) for b in range(self.batch_size)]
        if isinstance(parts[0][0], np.ndarray):
            stack = np.stack
        else:
            import torch
            stack = torch.stack
        image = stack([p[0] for p in parts])
        seg = stack([p[1] for p in parts])
        axis = 4 if self.data_format == 'NHWC' else 2
        if self.data_format not in ('NHWC', 'NCHW'):
            raise ValueError()
        image, seg = (image[..., None], seg[..., None]) if axis == 4 else (image[:, :, None], seg[:, :, None])
        full_seg = np.stack([p[2] for p in parts])
        is_last = np.stack([p[3][-1] for p in parts]).astype(np.float32)
        return image, seg, full_seg, is_last


class CTCInferenceReader:
    """Mirror of DataHandeling.CTCInferenceReader (DataHandeling.py:1572-1596): ``.dataset`` iterates the frames of a
    sequence folder -- the first ``pre_sequence_frames`` played in reverse, then all frames in order -- as float32 images,
    z-scored per frame.  Host I/O (OpenCV); an iterable instead of a tf.data.Dataset."""

    def __init__(self, data_path, filename_format='t*.tif', normalize=True, pre_sequence_frames=0):
        import glob
        import os
        file_list = glob.glob(os.path.join(data_path, filename_format))
        if len(file_list) == 0:
            raise ValueError('Could not read images from: {}'.format(os.path.join(data_path, filename_format)))
        file_list.sort()
        self.file_list = file_list[:pre_sequence_frames][::-1] + file_list
        self.normalize = normalize

    def _gen(self):
        import cv2
        for file in self.file_list:
            img = cv2.imread(file, -1)
            if img is None:
                raise ValueError('Could not read image: {}'.format(file))
            img = img.astype(np.float32)
            if self.normalize:
                img = (img - img.mean())
                img = img / (img.std())
            yield img

    @property
    def dataset(self):
        return self._gen()
