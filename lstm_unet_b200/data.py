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
