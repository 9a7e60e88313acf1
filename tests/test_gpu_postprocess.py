"""GPU parity tests of the instance-labelling post-processing (lu_postprocess through postprocess.PostProcessor) --
bit-exact against the golden vectors made by executing the reference's own statements, and against the pinned oracle
on seeded inputs at the benchmark frame size."""
import numpy as np
import pytest
import torch

from oracle import postprocess_oracle as P
from tests.test_postprocess_oracle import golden_cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', list(golden_cases()), ids=lambda c: c[0])
def test_matches_reference_vectors(case):
    from lstm_unet_b200.postprocess import PostProcessor
    name, sm, want, num, kw = case
    pp = PostProcessor(**kw)
    got = pp(sm).numpy()
    assert got.dtype == np.uint16 and got.shape == want.shape
    assert pp.info()[0, 0] == num and pp.info()[0, 1] == want.max()
    assert np.array_equal(got, want)


@pytest.mark.parametrize('kind,H,W,n', [('cells', 512, 512, 4), ('noise', 256, 320, 3), ('cells', 1024, 1024, 1),
                                        ('noise', 130, 70, 8), ('full', 300, 200, 1), ('empty', 64, 64, 2)])
def test_matches_oracle_large_and_batched(kind, H, W, n):
    from lstm_unet_b200.postprocess import PostProcessor
    sms = np.stack([P.synthetic_softmax(H, W, 50 + i, kind) for i in range(n)])
    kw = dict(edge_dist=3, min_cell_size=3, max_cell_size=(10 ** 7 if kind == 'full' else 200), FOV=2)
    want = np.stack([P.postprocess_frame(s, **kw) for s in sms])
    pp = PostProcessor(**kw)
    dev = torch.from_numpy(sms).cuda()
    got = pp(dev).numpy()
    assert np.array_equal(got, want)
    # the same frames again through the same workspace, and in the other layout: no state leaks between calls
    assert np.array_equal(pp(dev).numpy(), want)
    last = PostProcessor(data_format='NHWC', **kw)(dev.permute(0, 2, 3, 1).contiguous()).numpy()
    assert np.array_equal(last, want)


def test_sequential_pass_on_device():
    from lstm_unet_b200.postprocess import PostProcessor
    kw = dict(edge_dist=4, min_cell_size=1, max_cell_size=10000)
    flagged = 0
    for seed in range(8):
        sm = P.synthetic_softmax(48, 48, 100 + seed, 'noise')
        pp = PostProcessor(**kw)
        got = pp(sm).numpy()
        flagged += int(pp.info()[0, 2])
        assert np.array_equal(got, P.postprocess_frame(sm, **kw)), seed
    assert flagged > 0


def test_repeatable_under_contention():
    """the union-find and the atomics must give the same labels on every run (the roots are minima, not winners)"""
    from lstm_unet_b200.postprocess import PostProcessor
    sms = torch.from_numpy(np.stack([P.synthetic_softmax(384, 384, 70 + i, 'noise') for i in range(6)])).cuda()
    pp = PostProcessor(edge_dist=2, min_cell_size=2, max_cell_size=500)
    first = pp(sms).numpy().copy()
    for _ in range(5):
        assert np.array_equal(pp(sms).numpy(), first)


def test_model_softmax_to_labels_stays_on_device():
    """Inference2D's frame loop: model soft-max (device) -> labels; equals the oracle applied to the soft-max the
    model produced"""
    from lstm_unet_b200.Networks import ULSTMnet2D
    from lstm_unet_b200.postprocess import PostProcessor
    from oracle import lstm_unet_oracle as O
    net = {'down_conv_kernels': [[(3, 16), (3, 16)], [(3, 32), (3, 32)]], 'lstm_kernels': [[(5, 16)], [(5, 32)]],
           'up_conv_kernels': [[(3, 32), (3, 32)], [(3, 16), (3, 16), (1, 3)]]}
    params = O.init_params(net, seed=3, randomize_bn=True)
    m = ULSTMnet2D(net, 'NCHW', True, precision='bf16')
    m.set_weights_dict({k: v.numpy().copy() for k, v in params.items()})
    x = np.random.default_rng(0).standard_normal((2, 2, 1, 64, 72)).astype(np.float32) * 3
    _, softmax = m(x, training=False)
    pp = PostProcessor(edge_dist=2, min_cell_size=1, max_cell_size=1000)
    labels = pp(softmax).numpy()
    assert labels.shape == (2, 2, 64, 72)
    sm = softmax.numpy()
    for b in range(2):
        for t in range(2):
            assert np.array_equal(labels[b, t], P.postprocess_frame(sm[b, t], edge_dist=2, min_cell_size=1,
                                                                    max_cell_size=1000))


def test_train_export_then_inference_with_labels(tmp_path):
    """train2D.train exports model.ckpt (TF2 tensor bundle) + model_params.pickle (train2D.py:232-240); Inference2D.inference
    loads them, streams frames, labels every frame on the device and writes the uint16 masks (Inference2D.py:27-35,59-126)"""
    import cv2
    from lstm_unet_b200 import Params, train2D, Inference2D
    net = {'down_conv_kernels': [[(3, 16), (3, 16)], [(3, 32), (3, 32)]], 'lstm_kernels': [[(5, 16)], [(5, 32)]],
           'up_conv_kernels': [[(3, 32), (3, 32)], [(3, 16), (3, 16), (1, 3)]]}
    p = Params.CTCParams({'net_kernel_params': net, 'crop_size': (32, 32), 'batch_size': 2, 'unroll_len': 2, 'dry_run': False,
                          'learning_rate': 1e-3, 'validation_interval': 100, 'print_to_console_interval': 100,
                          'save_checkpoint_dir': str(tmp_path / 'ckpt'), 'save_log_dir': str(tmp_path / 'log')})
    train2D.params = p
    train2D.train(num_iterations=2, log=lambda *a: None)
    save_dir = p.experiment_save_dir
    import os
    assert os.path.exists(os.path.join(save_dir, 'model.ckpt.index')) and os.path.exists(os.path.join(save_dir, 'model_params.pickle'))
    frames = [np.random.default_rng(i).standard_normal((40, 48)).astype(np.float32) * 2 for i in range(3)]
    ip = Params.CTCInferenceParams({'model_path': save_dir, 'pre_sequence_frames': 1, 'dry_run': False, 'min_cell_size': 1,
                                    'max_cell_size': 10000, 'output_path': str(tmp_path / 'out'), 'save_intermediate': True})
    Inference2D.params = ip
    got = []
    outs = Inference2D.inference(frames, on_labels=lambda t, lab: got.append((t, lab.copy())))
    assert len(outs) == 3 and [t for t, _ in got] == [0, 1, 2]
    for (t, lab), sm in zip(got, outs):
        assert lab.dtype == np.uint16 and lab.shape == (40, 48)
        assert np.array_equal(lab, P.postprocess_frame(sm, edge_dist=2, min_cell_size=1, max_cell_size=10000))
        disk = cv2.imread(str(tmp_path / 'out' / ('mask%03d.tif' % t)), -1)
        assert disk is not None and np.array_equal(disk, lab)
    # the same through the sequence reader (Inference2D.py:42-43): TIFF frames on disk, z-scored by the reader
    seq_dir = tmp_path / 'seq'
    os.makedirs(seq_dir)
    raws = [np.random.default_rng(10 + i).integers(0, 3000, size=(40, 48)).astype(np.uint16) for i in range(3)]
    for i, a in enumerate(raws):
        cv2.imwrite(str(seq_dir / ('t%03d.tif' % i)), a)
    ip2 = Params.CTCInferenceParams({'model_path': save_dir, 'pre_sequence_frames': 2, 'dry_run': False, 'min_cell_size': 1,
                                     'max_cell_size': 10000, 'output_path': str(tmp_path / 'out2'), 'sequence_path': str(seq_dir)})
    Inference2D.params = ip2
    outs2 = Inference2D.inference()
    assert len(outs2) == 3 and len(Inference2D.last_labels) == 3
    for lab, sm in zip(Inference2D.last_labels, outs2):
        assert np.array_equal(lab, P.postprocess_frame(sm, edge_dist=2, min_cell_size=1, max_cell_size=10000))
