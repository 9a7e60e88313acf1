"""Generates tests/golden/postprocess.npz by EXECUTING the reference's own post-processing statements.

The instance-labelling code of the reference is inline in ``Inference2D.inference`` (Inference2D.py:64-123) and the
module imports TensorFlow at the top, so it cannot be imported; this script reads the file from /root/reference at
run time, takes the statements between the soft-max threshold and the TIFF write as they stand (nothing is copied into
this repository), and executes them -- with the reference's own ``bbox_crop`` / ``bbox_fill`` from utils.py -- on seeded
synthetic soft-max maps.  Inputs are stored as float32, outputs as uint16, plus the OpenCV / SciPy versions used.

Run in the build container only (``python tests/golden/make_postprocess_golden.py``); /root/reference does not exist on
the GPU box, the committed .npz does.
"""
import os
import sys
import textwrap
import types

import cv2
import numpy as np
import scipy
import scipy.ndimage  # noqa: F401  (the reference statements use scipy.ndimage.morphology.*)

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = os.environ.get('LSTM_UNET_REFERENCE', '/root/reference')


def _reference_statements():
    src = open(os.path.join(REF, 'Inference2D.py')).read().split('\n')
    first = next(i for i, l in enumerate(src) if 'seg_edge = np.greater_equal(image_softmax_np[2]' in l)
    last = next(i for i, l in enumerate(src) if 'out_fname = base_out_fname.format(time=t)' in l)
    body = textwrap.dedent('\n'.join(src[first:last]))
    utils_src = open(os.path.join(REF, 'utils.py')).read()
    helpers = {'np': np}
    exec(compile(utils_src[utils_src.index('def bbox_crop'):], 'utils.py', 'exec'), helpers)
    return compile(body, 'Inference2D.py[%d:%d]' % (first + 1, last), 'exec'), helpers


_CODE, _HELPERS = _reference_statements()


def reference_postprocess(softmax, edge_dist=2, min_cell_size=10, max_cell_size=100, FOV=0):
    params = types.SimpleNamespace(edge_dist=edge_dist, min_cell_size=min_cell_size, max_cell_size=max_cell_size,
                                   FOV=FOV, save_intermediate=False, data_format='NCHW')
    ns = {'np': np, 'scipy': scipy, 'cv2': cv2, 'bbox_crop': _HELPERS['bbox_crop'], 'bbox_fill': _HELPERS['bbox_fill'],
          'params': params, 'image_softmax_np': np.array(softmax, dtype=np.float32, copy=True)}
    exec(_CODE, ns)
    return ns['labels_out'], int(ns['num_cells'])


CASES = [
    # name, H, W, kind, seed, params
    ('cells_96x128', 96, 128, 'cells', 1, {}),
    ('cells_64x64_fov', 64, 64, 'cells', 2, {'FOV': 6}),
    ('cells_80x72_dist3', 80, 72, 'cells', 3, {'edge_dist': 3, 'max_cell_size': 400}),
    ('cells_50x61_dist2p5', 50, 61, 'cells', 4, {'edge_dist': 2.5, 'min_cell_size': 4}),
    ('noise_48x56', 48, 56, 'noise', 5, {'min_cell_size': 1}),
    ('noise_40x40_fov', 40, 40, 'noise', 6, {'min_cell_size': 2, 'FOV': 3}),
    ('noise_33x47_dist4', 33, 47, 'noise', 7, {'min_cell_size': 1, 'edge_dist': 4}),
    ('empty_24x24', 24, 24, 'empty', 8, {}),
    ('full_24x32', 24, 32, 'full', 9, {'max_cell_size': 10000}),
    ('cells_200x200', 200, 200, 'cells', 10, {'max_cell_size': 300}),
    ('cells_1x17', 1, 17, 'noise', 11, {'min_cell_size': 1}),
    ('cells_9x2', 9, 2, 'noise', 12, {'min_cell_size': 1}),
    # larger frames, soft-max stored as float16 (the reference is fed the float16-rounded values)
    ('cells_256x256_f16', 256, 256, 'cells', 13, {'max_cell_size': 200}),
    ('noise_128x160_f16', 128, 160, 'noise', 14, {'min_cell_size': 2, 'edge_dist': 3, 'FOV': 5}),
]


def main():
    from oracle.postprocess_oracle import synthetic_softmax
    out = {'versions': np.array(['cv2 ' + cv2.__version__, 'scipy ' + scipy.__version__, 'numpy ' + np.__version__])}
    names = []
    for name, H, W, kind, seed, kw in CASES:
        sm = synthetic_softmax(H, W, seed, kind)
        if name.endswith('_f16'):
            sm = sm.astype(np.float16).astype(np.float32)
        labels, num = reference_postprocess(sm, **kw)
        names.append(name)
        out[name + '/softmax'] = sm.astype(np.float16) if name.endswith('_f16') else sm
        out[name + '/labels'] = labels.astype(np.uint16)
        out[name + '/num_cells'] = np.int64(num)
        out[name + '/params'] = np.array([kw.get('edge_dist', 2), kw.get('min_cell_size', 10),
                                          kw.get('max_cell_size', 100), kw.get('FOV', 0)], dtype=np.float64)
        print('%-22s cc=%4d kept=%4d' % (name, num, int(labels.max())))
    out['names'] = np.array(names)
    np.savez_compressed(os.path.join(HERE, 'postprocess.npz'), **out)


if __name__ == '__main__':
    main()
