"""The SEG-measure oracle (oracle/seg_oracle.py) against the vectors produced by the reference's own ``seg_numpy``
(tests/golden/make_seg_golden.py)."""
import os

import numpy as np
import pytest

from oracle import seg_oracle as S

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'seg.npz')


def seg_cases():
    g = np.load(GOLD)
    for name in g['names']:
        name = str(name)
        yield name, g[name + '/labels'].astype(np.float32), g[name + '/logits'].astype(np.float32), float(g[name + '/seg'])


@pytest.mark.parametrize('case', list(seg_cases()), ids=lambda c: c[0])
def test_seg_oracle_matches_reference_vectors(case):
    name, labels, logits, want = case
    got = S.seg_measure(labels, logits)
    assert got == pytest.approx(want, rel=1e-6, abs=1e-7)


def test_no_objects_is_nan():
    z = np.zeros((1, 1, 1, 8, 8), np.float32)
    lg = np.zeros((1, 1, 3, 8, 8), np.float32)
    assert np.isnan(S.seg_measure(z, lg)) and np.isnan(float(np.load(GOLD)['empty/seg']))


def test_accuracy_counts_ignore_labels_as_wrong():
    lab = np.array([-1, 0, 1, 2], np.float32).reshape(1, 1, 1, 2, 2)
    lg = np.zeros((1, 1, 3, 2, 2), np.float32)
    lg[0, 0, 0] = 1          # predicts class 0 everywhere
    assert S.accuracy(lab, lg) == 0.25
