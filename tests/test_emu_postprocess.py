"""CPU tests of the post-processing host logic and kernels' arithmetic: the TEST-ONLY host build of the library
(tests/_emu, -DLU_HOST_EMU: every kernel body runs as plain loops, one "thread" per CTA) through the same C-ABI and the
same PostProcessor driver the product uses, against the reference-pinned oracle and the golden vectors."""
import numpy as np
import pytest

from oracle import postprocess_oracle as P
from tests.emu_backend import NumpyBackend, build_emu
from tests.test_postprocess_oracle import golden_cases


def emu_post(**kw):
    from lstm_unet_b200 import _lib
    from lstm_unet_b200.postprocess import PostProcessor
    lib = _lib.load_library(build_emu())
    assert lib.lu_is_cuda_build() == 0
    return PostProcessor(_lib_override=lib, _backend=NumpyBackend(), **kw)


@pytest.mark.parametrize('case', list(golden_cases()), ids=lambda c: c[0])
def test_emu_matches_reference_vectors(case):
    name, sm, want, num, kw = case
    pp = emu_post(**kw)
    got = pp(sm)
    info = pp.info()
    assert got.dtype == np.uint16 and got.shape == want.shape
    assert info[0, 0] == num
    assert info[0, 1] == want.max()
    assert np.array_equal(got, want)


@pytest.mark.parametrize('kind,seed,H,W', [('noise', 21, 37, 53), ('noise', 22, 64, 64), ('cells', 23, 120, 90),
                                           ('cells', 24, 65, 33), ('noise', 25, 3, 3), ('noise', 26, 2, 70)])
def test_emu_matches_oracle_batched(kind, seed, H, W):
    """several frames in one call, per-frame independence, both layouts"""
    sms = np.stack([P.synthetic_softmax(H, W, seed * 10 + i, kind) for i in range(3)])
    kw = dict(edge_dist=3, min_cell_size=2, max_cell_size=150, FOV=1 if min(H, W) > 4 else 0)
    want = np.stack([P.postprocess_frame(s, **kw) for s in sms])
    got = emu_post(**kw)(sms)
    assert np.array_equal(got, want)
    got_last = emu_post(data_format='NHWC', **kw)(np.ascontiguousarray(sms.transpose(0, 2, 3, 1)))
    assert np.array_equal(got_last, want)


def test_emu_sequential_pass_is_exercised_and_exact():
    """an edge ring around a cell that itself encloses another cell: the enclosed cell's pixels already carry a label,
    so the reference's sequential `labels += holes * n` is not separable -- the device path must notice and redo it"""
    H = W = 40
    z = np.zeros((3, H, W), np.float32)
    z[0] = 5
    yy, xx = np.mgrid[0:H, 0:W]
    r = np.sqrt((yy - 20) ** 2 + (xx - 20) ** 2)
    z[1][(r >= 9) & (r < 12)] = 9          # ring-shaped cell (label 1) -- its hole is filled by the global fill
    z[1][r < 3] = 9
    sm = np.exp(z) / np.exp(z).sum(0)
    # global fill makes it one disc; build the nested case after the fill instead: ring of EDGE pixels around cell A
    # that are nearest to cell B
    z = np.zeros((3, H, W), np.float32)
    z[0] = 5
    z[1][(np.abs(yy - 20) <= 2) & (np.abs(xx - 20) <= 2)] = 9                    # inner cell
    z[1][(xx >= 30) & (xx <= 33) & (yy >= 5) & (yy <= 35)] = 9                   # outer cell (a bar on the right)
    ring = (np.maximum(np.abs(yy - 20), np.abs(xx - 20)) == 9)
    z[2][ring] = 9                                                               # edge ring, closest to the bar? no:
    sm = (np.exp(z) / np.exp(z).sum(0)).astype(np.float32)
    for e in (2, 12):
        kw = dict(edge_dist=e, min_cell_size=1, max_cell_size=10000)
        pp = emu_post(**kw)
        got = pp(sm)
        assert np.array_equal(got, P.postprocess_frame(sm, **kw))
    # adversarial noise: nested labels inside holes do occur; at least one of these frames needs the sequential pass
    flagged = 0
    for seed in range(8):
        sm = P.synthetic_softmax(48, 48, 100 + seed, 'noise')
        kw = dict(edge_dist=4, min_cell_size=1, max_cell_size=10000)
        pp = emu_post(**kw)
        got = pp(sm)
        flagged += int(pp.info()[0, 2])
        assert np.array_equal(got, P.postprocess_frame(sm, **kw)), seed
    assert flagged > 0


def test_edge_d2_limit_follows_float64_sqrt():
    from lstm_unet_b200.postprocess import edge_d2_limit
    for e in (0, 0.5, 1, 1.5, 2, 2.5, 3, 7, 10.3, np.sqrt(5.0), 2.0000001):
        assert edge_d2_limit(e) == P.edge_dist_threshold(e), e


def test_bad_arguments_raise():
    from lstm_unet_b200.session import LuError
    with pytest.raises(ValueError):
        emu_post()(np.zeros((4, 8, 8), np.float32))
    with pytest.raises(LuError):
        emu_post(FOV=9)(np.zeros((3, 8, 8), np.float32))


def test_emu_randomized_sweep_against_oracle():
    """seeded sweep over frame shapes, map statistics and parameters (incl. non-integer edge distances, FOV, degenerate
    1-pixel-wide frames): every label image equals the oracle's"""
    rng = np.random.default_rng(2024)
    flagged = 0
    for trial in range(60):
        H, W = int(rng.integers(1, 60)), int(rng.integers(2, 60))
        kind = ('noise', 'cells')[trial % 2] if min(H, W) >= 8 else 'noise'
        sm = P.synthetic_softmax(H, W, 5000 + trial, kind)
        if trial % 7 == 0:                       # exact ties in the arg-max and at the edge threshold
            sm = np.round(sm * 4) / 4
            sm = (sm / np.maximum(sm.sum(0, keepdims=True), 1e-6)).astype(np.float32)
        kw = dict(edge_dist=float(rng.choice([0, 1, 1.5, 2, 2.5, 3, 5])), min_cell_size=int(rng.integers(0, 6)),
                  max_cell_size=int(rng.choice([20, 100, 10 ** 6])), FOV=int(rng.integers(0, max(1, min(W - 1, 4)))))
        pp = emu_post(**kw)
        got = pp(sm)
        flagged += int(pp.info()[0, 2])
        assert np.array_equal(got, P.postprocess_frame(sm, **kw)), (trial, H, W, kind, kw)
    assert flagged > 0


def test_emu_nan_and_inf_soft_max_follow_numpy():
    """np.argmax treats the first NaN as the maximum and `NaN >= 0.2` is False; +inf wins like any large value"""
    rng = np.random.default_rng(9)
    sm = P.synthetic_softmax(24, 28, 77, 'noise')
    for c in range(3):
        m = rng.random((24, 28)) < 0.05
        sm[c][m] = np.nan
    sm[1][rng.random((24, 28)) < 0.03] = np.inf
    sm[2][rng.random((24, 28)) < 0.03] = -np.inf
    kw = dict(edge_dist=2, min_cell_size=1, max_cell_size=10 ** 6)
    assert np.array_equal(emu_post(**kw)(sm), P.postprocess_frame(sm, **kw))
