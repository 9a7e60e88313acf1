"""CPU tests of the device SEG measure / accuracy (lu_seg_measure) through the TEST-ONLY host build: against the vectors
produced by the reference's own ``seg_numpy`` and against the oracle on seeded inputs, both layouts."""
import numpy as np
import pytest

from oracle import seg_oracle as S
from tests.emu_backend import NumpyBackend, build_emu
from tests.test_seg_oracle import seg_cases


def emu_metrics(channel_axis=2):
    from lstm_unet_b200 import _lib, losses
    lib = _lib.load_library(build_emu())
    assert lib.lu_is_cuda_build() == 0
    return losses.seg_measure(channel_axis, _lib_override=lib, _backend=NumpyBackend())


@pytest.mark.parametrize('case', list(seg_cases()), ids=lambda c: c[0])
def test_emu_seg_matches_reference_vectors(case):
    name, labels, logits, want = case
    calc = emu_metrics()
    assert calc(labels, logits) == pytest.approx(want, rel=1e-6, abs=1e-7)
    assert calc.last_accuracy == pytest.approx(S.accuracy(labels, logits), rel=1e-12)


@pytest.mark.parametrize('kind,seed,shape', [('blobs', 11, (2, 3, 40, 52)), ('noise', 12, (1, 2, 33, 65)),
                                             ('blobs', 13, (3, 1, 7, 100)), ('noise', 14, (1, 1, 2, 2))])
def test_emu_seg_matches_oracle_both_layouts(kind, seed, shape):
    B, T, H, W = shape
    labels, logits = S.synthetic_pair(B, T, H, W, seed, kind)
    want = S.seg_measure(labels, logits)
    got = emu_metrics()(labels, logits)
    assert (np.isnan(want) and np.isnan(got)) or got == pytest.approx(want, rel=1e-6, abs=1e-7)
    last = emu_metrics(4)(labels.transpose(0, 1, 3, 4, 2).copy(), logits.transpose(0, 1, 3, 4, 2).copy())
    assert (np.isnan(want) and np.isnan(last)) or last == pytest.approx(want, rel=1e-6, abs=1e-7)


def test_emu_seg_nan_without_objects_and_bad_shapes():
    calc = emu_metrics()
    lab = np.zeros((1, 1, 1, 8, 8), np.float32)
    lg = np.zeros((1, 1, 3, 8, 8), np.float32)
    assert np.isnan(calc(lab, lg)) and calc.last_accuracy == 1.0
    with pytest.raises(ValueError):
        calc(lab[..., :4], lg)
    from lstm_unet_b200 import losses
    with pytest.raises(ValueError):
        losses.seg_measure(2, three_d=True)


def test_emu_seg_randomized_sweep():
    rng = np.random.default_rng(77)
    for trial in range(25):
        B, T = int(rng.integers(1, 3)), int(rng.integers(1, 4))
        H, W = int(rng.integers(1, 50)), int(rng.integers(1, 50))
        labels, logits = S.synthetic_pair(B, T, H, W, 900 + trial, ('noise', 'blobs')[trial % 2] if min(H, W) > 6 else 'noise')
        if trial % 5 == 0:
            logits = np.round(logits)              # arg-max ties
        want, acc = S.seg_measure(labels, logits), S.accuracy(labels, logits)
        calc = emu_metrics()
        got = calc(labels, logits)
        assert (np.isnan(want) and np.isnan(got)) or got == pytest.approx(want, rel=1e-6, abs=1e-7), (trial, B, T, H, W)
        assert calc.last_accuracy == pytest.approx(acc, rel=1e-12, abs=1e-15)
