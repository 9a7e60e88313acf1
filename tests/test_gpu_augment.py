"""GPU parity tests of the device augmentation (lu_augment_sequence / lu_elastic_coords through
augment.SequenceAugmenter and data.CTCRAMReaderSequence2D): against the vectors made by the reference's own helpers and
against the pinned oracle at a training crop size.  Segmentations bit-exact; images to float32 rounding of the frame
mean (rtol 2e-6 + atol 2e-4 on values of order 1e3); elastic coordinates to 1e-10 pixel."""
import numpy as np
import pytest
import torch

from oracle import augment_oracle as A
from tests.test_augment_oracle import augment_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', list(augment_cases()), ids=lambda c: c[0])
def test_matches_reference_vectors(case):
    from lstm_unet_b200.augment import SequenceAugmenter
    name, c = case
    aug = SequenceAugmenter()
    flip, rot = (int(c['flip_rot'][0]), int(c['flip_rot'][1])), int(c['flip_rot'][2])
    coords = None
    if 'rand2' in c:
        H, W = c['img'].shape[1:]
        coords = aug.elastic_coords(c['rand2'], W * 2, W * 0.15)
        np.testing.assert_allclose(coords.cpu().numpy().reshape(2, H, W), c['coords'], rtol=0, atol=1e-10)
    img, seg = aug.augment(c['img'], c['seg'], c['contrast'], c['brightness'], c.get('affine'), coords, flip, rot)
    assert tuple(img.shape) == c['out_img'].shape
    assert np.array_equal(seg.cpu().numpy(), c['out_seg'])
    np.testing.assert_allclose(img.cpu().numpy(), c['out_img'], rtol=2e-6, atol=2e-4)


def test_matches_oracle_at_training_crop():
    from lstm_unet_b200.augment import SequenceAugmenter, random_affine
    T, H, W = 3, 160, 160
    imgs, segs = A.synthetic_sequence(T, H, W, 41, unlabeled_every=3)
    rs = np.random.RandomState(2)
    affine = random_affine((H, W), W * 0.08, rs)
    rand2 = np.stack([rs.rand(H, W), rs.rand(H, W)])
    contrast = (rs.rand(T) + 0.5).astype(np.float32)
    brightness = ((rs.rand(T) - 0.5) * 0.2 * imgs.max()).astype(np.float32)
    aug = SequenceAugmenter()
    coords = aug.elastic_coords(rand2, W * 2, W * 0.15)
    ref_coords = A.elastic_coords(rand2, W * 2, W * 0.15)
    np.testing.assert_allclose(coords.cpu().numpy().reshape(2, H, W), ref_coords, rtol=0, atol=1e-9)
    img, seg = aug.augment(torch.from_numpy(imgs).cuda(), torch.from_numpy(segs).cuda(), contrast, brightness, affine, coords,
                           (1, 1), 3)
    img, seg = img.cpu().numpy(), seg.cpu().numpy()
    for t in range(T):
        ri, rs_ = A.augment_frame(imgs[t], segs[t], contrast[t], brightness[t], affine, ref_coords, (1, 1), 3)
        assert np.array_equal(seg[t], rs_)
        np.testing.assert_allclose(img[t], ri, rtol=2e-6, atol=2e-4)
    assert np.all(seg[2] == -1)                       # the frame without annotation passes through


def test_reader_feeds_the_model_on_the_device():
    """train2D's loop with the real reader mirror: batches are device tensors in the model's layout"""
    from lstm_unet_b200.data import CTCRAMReaderSequence2D
    from lstm_unet_b200.Networks import ULSTMnet2D, Adam
    seqs = []
    for s in range(2):
        imgs, segs = A.synthetic_sequence(8, 48, 48, 60 + s)
        seqs.append({'images': (imgs - imgs.mean()) / imgs.std(), 'segs': segs, 'full_seg': np.ones(8)})
    rd = CTCRAMReaderSequence2D(sequences=seqs, image_crop_size=(32, 32), unroll_len=2, batch_size=2, seed=0, elastic_seed=1)
    rd.start_queues()
    net = {'down_conv_kernels': [[(3, 16), (3, 16)], [(3, 32), (3, 32)]], 'lstm_kernels': [[(5, 16)], [(5, 32)]],
           'up_conv_kernels': [[(3, 32), (3, 32)], [(3, 16), (3, 16), (1, 3)]]}
    model = ULSTMnet2D(net, 'NCHW', False, train=True)
    opt = Adam(lr=1e-3)
    for _ in range(3):
        image, seg, full, is_last = rd.get_batch()
        assert image.is_cuda and tuple(image.shape) == (2, 2, 1, 32, 32) and tuple(seg.shape) == (2, 2, 1, 32, 32)
        _, _, loss = model.train_step(image, seg, [0.15, 0.25, 0.6], opt)
        model.reset_states_per_batch(is_last)
        assert np.isfinite(float(loss))
