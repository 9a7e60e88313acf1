"""Generates tests/golden/augment.npz by EXECUTING the reference's own augmentation helpers.

DataHandeling.py imports TensorFlow (queues) and utils (-> Networks -> Keras) at module level, but the arithmetic of
the training reader's augmentation lives in static methods of ``CTCRAMReaderSequence2D`` that only use numpy, OpenCV and
SciPy.  This script imports the module from /root/reference with permissive stand-ins registered as ``tensorflow`` and
``utils`` and then replays, statement for statement, the per-frame part of ``_load_and_enqueue``
(DataHandeling.py:330-377): the statements themselves are read from the file and executed (np.random seeded, so the
contrast / brightness factors they draw are known), calling the reference's static helpers.  Nothing is copied into this repository.  Run in the build container only."""
import os
import sys
import types

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = os.environ.get('LSTM_UNET_REFERENCE', '/root/reference')


class _Anything(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Anything(k)

    def __call__(self, *a, **k):
        return _Anything('call')

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def reference_reader():
    for m in ('tensorflow', 'utils'):
        sys.modules[m] = _Anything(m)
    sys.path.insert(0, REF)
    try:
        import DataHandeling
    finally:
        sys.path.remove(REF)
        for m in ('tensorflow', 'utils'):
            del sys.modules[m]
    return DataHandeling.CTCRAMReaderSequence2D


def _frame_statements():
    """the per-frame statements of _load_and_enqueue (from the contrast draw to the rot90), compiled as they stand"""
    src = open(os.path.join(REF, 'DataHandeling.py')).read().split('\n')
    first = next(i for i, l in enumerate(src) if l.strip() == '# contrast factor between [0.5, 1.5]') - 1
    assert src[first].strip() == 'if self.randomize:'
    last = next(i for i in range(first, len(src)) if src[i].strip() == 'if self.return_dist:')
    import textwrap
    return compile(textwrap.dedent('\n'.join(src[first:last])), 'DataHandeling.py[%d:%d]' % (first + 1, last), 'exec')


_FRAME_CODE = None


def reference_frame(R, img_crop, seg_crop, np_seed, img_max, affine_matrix, indices, flip, rotate, randomize=True):
    """Executes the reference's own frame-loop statements (DataHandeling.py:330-377).  The contrast / brightness factors
    are drawn by those statements from np.random, seeded here; returns them with the augmented frame."""
    global _FRAME_CODE
    if _FRAME_CODE is None:
        _FRAME_CODE = _frame_statements()
    self = types.SimpleNamespace(randomize=randomize, elastic_augmentation=affine_matrix is not None,
                                 _adjust_contrast_=R._adjust_contrast_, _adjust_brightness_=R._adjust_brightness_,
                                 _get_transformed_image_=R._get_transformed_image_,
                                 _fix_transformed_segmentation=R._fix_transformed_segmentation)
    np.random.seed(np_seed)
    ns = {'np': np, 'cv2': cv2, 'self': self, 'img_crop': img_crop.copy(), 'seg_crop': seg_crop.copy(), 'img_max': img_max,
          'affine_matrix': affine_matrix, 'indices': indices, 'flip': flip, 'rotate': rotate, 'file_idx': 0,
          'sequence_folder': 'synthetic'}
    exec(_FRAME_CODE, ns)
    np.random.seed(np_seed)                      # the two draws the statements made, in their order
    contrast = np.random.rand() + 0.5
    brightness = (np.random.rand() - 0.5) * 0.2 * img_max
    return np.ascontiguousarray(ns['img_crop']), np.ascontiguousarray(ns['seg_crop']), contrast, brightness


CASES = [  # name, T, H, W, seed, elastic, flip, rot
    ('elastic_48', 3, 48, 48, 1, True, (0, 0), 0),
    ('elastic_flip_rot_40', 2, 40, 40, 2, True, (1, 1), 3),
    ('plain_flip_32x44', 2, 32, 44, 3, False, (1, 0), 0),
    ('plain_rot_36', 2, 36, 36, 4, False, (0, 1), 1),
    ('elastic_unlabeled_56', 4, 56, 56, 5, True, (0, 1), 2),
    ('elastic_64x40', 2, 64, 40, 6, True, (1, 0), 2),
]


def main():
    from oracle.augment_oracle import synthetic_sequence
    R = reference_reader()
    out, names = {}, []
    for name, T, H, W, seed, elastic, flip, rot in CASES:
        imgs, segs = synthetic_sequence(T, H, W, seed, unlabeled_every=3 if 'unlabeled' in name else 0)
        rng = np.random.RandomState(seed)
        if elastic:
            np.random.seed(seed)
            # _get_elastic_affine_matrix_ seeds its own RandomState(None): replay its statements with a seeded state
            src = open(os.path.join(REF, 'DataHandeling.py')).read()
            body = src[src.index('    def _get_elastic_affine_matrix_'):src.index('    @staticmethod\n    def _get_transformed_image_')]
            body = body.replace('np.random.RandomState(None)', 'np.random.RandomState(%d)' % seed)
            ns = {'np': np, 'cv2': cv2}
            exec(compile('if 1:\n' + body, 'DataHandeling.py[_get_elastic_affine_matrix_]', 'exec'), ns)
            affine, state = ns['_get_elastic_affine_matrix_']((H, W), W * 0.08)
            rs0 = np.random.RandomState(seed)
            rs0.uniform(-1, 1, size=(3, 2))                       # the draw the affine step consumed
            rand2 = np.stack([rs0.rand(H, W), rs0.rand(H, W)])    # what _get_indices4elastic_transform will draw
            indices = R._get_indices4elastic_transform((H, W), W * 2, W * 0.15, state)
            out[name + '/affine'] = affine
            out[name + '/rand2'] = rand2
            out[name + '/coords'] = np.stack([indices[0].reshape(H, W), indices[1].reshape(H, W)])
        else:
            affine = indices = None
        img_max = imgs.max()
        res = [reference_frame(R, imgs[t], segs[t], 1000 * seed + t, img_max, affine, indices, flip, rot) for t in range(T)]
        contrast = np.array([r[2] for r in res], np.float32)
        brightness = np.array([r[3] for r in res], np.float32)
        names.append(name)
        out[name + '/img'] = imgs
        out[name + '/seg'] = segs.astype(np.float32)
        out[name + '/contrast'] = contrast
        out[name + '/brightness'] = brightness
        out[name + '/flip_rot'] = np.array([flip[0], flip[1], rot], np.int64)
        out[name + '/out_img'] = np.stack([r[0] for r in res])
        out[name + '/out_seg'] = np.stack([r[1] for r in res]).astype(np.float32)
        print('%-24s img %s seg classes %s' % (name, res[0][0].shape, np.unique(out[name + '/out_seg'])))
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'augment.npz'), **out)


if __name__ == '__main__':
    main()
