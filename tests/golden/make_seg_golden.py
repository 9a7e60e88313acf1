"""Generates tests/golden/seg.npz by EXECUTING the reference's own SEG-measure arithmetic.

/root/reference/losses.py imports TensorFlow at module level but ``seg_numpy`` -- the numpy / SciPy closure inside
``seg_measure`` that does all the arithmetic (losses.py:40-71) -- does not use it.  This script imports the module from
/root/reference with an empty stand-in registered under the name ``tensorflow``, takes ``seg_numpy`` out of the closure
of the function ``seg_measure`` returns and runs it, as it stands, on seeded foreground masks.  Nothing is copied into
this repository.  Run in the build container only."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = os.environ.get('LSTM_UNET_REFERENCE', '/root/reference')


def reference_seg_numpy():
    sys.modules.setdefault('tensorflow', types.ModuleType('tensorflow'))
    sys.path.insert(0, REF)
    try:
        import losses as ref_losses
    finally:
        sys.path.remove(REF)
    calc = ref_losses.seg_measure(channel_axis=2, three_d=False, foreground_class_index=1)
    return next(c.cell_contents for c in calc.__closure__
                if callable(c.cell_contents) and c.cell_contents.__name__ == 'seg_numpy')


class _Tensor:            # seg_numpy calls .numpy() on its arguments
    def __init__(self, a):
        self.a = a

    def numpy(self):
        return self.a


CASES = [('blobs_2x2x48x56', 2, 2, 48, 56, 1, 'blobs'), ('blobs_1x3x64x64', 1, 3, 64, 64, 2, 'blobs'),
         ('noise_2x1x24x31', 2, 1, 24, 31, 3, 'noise'), ('noise_1x1x40x40', 1, 1, 40, 40, 4, 'noise'),
         ('blobs_1x1x128x128', 1, 1, 128, 128, 5, 'blobs'), ('noise_1x2x1x9', 1, 2, 1, 9, 6, 'noise')]


def main():
    from oracle.seg_oracle import foregrounds, synthetic_pair
    seg_numpy = reference_seg_numpy()
    out, names = {}, []
    for name, B, T, H, W, seed, kind in CASES:
        labels, logits = synthetic_pair(B, T, H, W, seed, kind)
        logits = logits.astype(np.float16).astype(np.float32)
        gt_fg, out_fg = foregrounds(labels, logits)
        v = seg_numpy(_Tensor(gt_fg), _Tensor(out_fg))
        names.append(name)
        out[name + '/labels'] = labels.astype(np.int8)
        out[name + '/logits'] = logits.astype(np.float16)     # stored as fp16: tests upcast, the values are the input
        out[name + '/seg'] = np.float64(v)
        print('%-20s SEG = %.9f (%s)' % (name, v, type(v).__name__))
    # no ground-truth object at all -> NaN
    z = np.zeros((1, 1, 8, 8), bool)
    out['empty/seg'] = np.float64(seg_numpy(_Tensor(z), _Tensor(z)))
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'seg.npz'), **out)


if __name__ == '__main__':
    main()
