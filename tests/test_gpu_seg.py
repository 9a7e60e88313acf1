"""GPU parity tests of the device SEG measure / accuracy (lu_seg_measure through losses.seg_measure): against the
vectors produced by the reference's own ``seg_numpy`` and against the pinned oracle at the training frame size.
Tolerance: the per-object IoUs are the same float32 values as the reference's; only the final mean is accumulated in a
different order (float64 atomics) -> 1e-6 relative."""
import numpy as np
import pytest
import torch

from oracle import seg_oracle as S
from tests.test_seg_oracle import seg_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', list(seg_cases()), ids=lambda c: c[0])
def test_seg_matches_reference_vectors(case):
    from lstm_unet_b200 import losses
    name, labels, logits, want = case
    calc = losses.seg_measure(2)
    assert calc(labels, logits) == pytest.approx(want, rel=1e-6, abs=1e-7)
    assert calc.last_accuracy == pytest.approx(S.accuracy(labels, logits), rel=1e-12)


@pytest.mark.parametrize('kind,shape', [('blobs', (4, 8, 256, 256)), ('noise', (2, 2, 200, 333)), ('blobs', (1, 1, 512, 512))])
def test_seg_matches_oracle_large(kind, shape):
    from lstm_unet_b200 import losses
    B, T, H, W = shape
    labels, logits = S.synthetic_pair(B, T, H, W, 31, kind)
    want, acc = S.seg_measure(labels, logits), S.accuracy(labels, logits)
    calc = losses.seg_measure(2)
    lab_d, lg_d = torch.from_numpy(labels).cuda(), torch.from_numpy(logits).cuda()
    for _ in range(3):                     # repeatable: the hash table and the union-find leave no state behind
        assert calc(lab_d, lg_d) == pytest.approx(want, rel=1e-6, abs=1e-7)
        assert calc.last_accuracy == pytest.approx(acc, rel=1e-12)
    last = losses.seg_measure(4)
    got = last(lab_d.permute(0, 1, 3, 4, 2).contiguous(), lg_d.permute(0, 1, 3, 4, 2).contiguous())
    assert got == pytest.approx(want, rel=1e-6, abs=1e-7)


def test_train_loop_reports_metrics():
    """train2D.train: every step records the SEG measure and the accuracy of that step's logits (train2D.py:97-102)"""
    from lstm_unet_b200 import Params, train2D
    net = {'down_conv_kernels': [[(3, 16), (3, 16)], [(3, 32), (3, 32)]], 'lstm_kernels': [[(5, 16)], [(5, 32)]],
           'up_conv_kernels': [[(3, 32), (3, 32)], [(3, 16), (3, 16), (1, 3)]]}
    p = Params.CTCParams({'net_kernel_params': net, 'crop_size': (32, 32), 'batch_size': 2, 'unroll_len': 2,
                          'learning_rate': 1e-3, 'validation_interval': 2, 'print_to_console_interval': 100})
    train2D.params = p
    losses_seen = train2D.train(num_iterations=4, log=lambda *a: None)
    m = train2D.train.metrics
    assert len(losses_seen) == 4 and len(m['train']['SEG']) == 4 and len(m['val']['SEG']) == 2
    assert all(0.0 <= a <= 1.0 for a in m['train']['accuracy'])
    assert all(np.isnan(v) or 0.0 <= v <= 1.0 for v in m['train']['SEG'])
