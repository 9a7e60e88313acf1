"""Generates tests/golden/augment.npz by EXECUTING the reference's own augmentation helpers.

DataHandeling.py imports TensorFlow (queues) and utils (-> Networks -> Keras) at module level, but the arithmetic of
the training reader's augmentation lives in static methods of ``CTCRAMReaderSequence2D`` that only use numpy, OpenCV and
SciPy.  This script imports the module from /root/reference with permissive stand-ins registered as ``tensorflow`` and
``utils`` and then replays, statement for statement, the per-frame part of ``_load_and_enqueue``
(DataHandeling.py:330-380) by calling those static methods as they stand -- with explicit, seeded parameters in place of
the reader's np.random draws.  Nothing is copied into this repository.  Run in the build container only."""
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


def reference_frame(R, img_crop, seg_crop, contrast, brightness, affine_matrix, indices, flip, rotate, randomize=True):
    """the body of the frame loop of _load_and_enqueue, calling the reference's helpers"""
    img_crop, seg_crop = img_crop.copy(), seg_crop.copy()
    if randomize:
        img_crop = R._adjust_contrast_(img_crop, contrast)
        img_crop = R._adjust_brightness_(img_crop, brightness)
    if affine_matrix is not None:
        img_crop = R._get_transformed_image_(img_crop, affine_matrix, indices)
        if not np.equal(seg_crop, -1).all():
            seg_not_valid = np.equal(seg_crop, -1)
            labeled_gt = seg_crop
            labeled_gt[:, 0] = 0
            labeled_gt[:, -1] = 0
            labeled_gt[-1, :] = 0
            labeled_gt[0, :] = 0
            trans_seg = R._get_transformed_image_(labeled_gt.astype(np.float32), affine_matrix, indices, seg=True)
            trans_not_valid = R._get_transformed_image_(seg_not_valid.astype(np.float32), affine_matrix, indices, seg=True)
            trans_seg_fix = R._fix_transformed_segmentation(trans_seg)
            trans_not_valid = np.logical_or(np.greater(trans_not_valid, 0.5), np.equal(trans_seg, -1))
            seg_crop = trans_seg_fix
            seg_crop[trans_not_valid] = -1
    else:
        seg_crop = R._fix_transformed_segmentation(seg_crop)
    if flip[0]:
        img_crop, seg_crop = cv2.flip(img_crop, 0), cv2.flip(seg_crop, 0)
    if flip[1]:
        img_crop, seg_crop = cv2.flip(img_crop, 1), cv2.flip(seg_crop, 1)
    if rotate > 0:
        img_crop, seg_crop = np.rot90(img_crop, rotate), np.rot90(seg_crop, rotate)
    return np.ascontiguousarray(img_crop), np.ascontiguousarray(seg_crop)


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
        contrast = (rng.rand(T) + 0.5).astype(np.float32)
        brightness = ((rng.rand(T) - 0.5) * 0.2 * imgs.max()).astype(np.float32)
        res = [reference_frame(R, imgs[t], segs[t], contrast[t], brightness[t], affine, indices, flip, rot) for t in range(T)]
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
