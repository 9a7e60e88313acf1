"""N>1 path on CPU: two gloo ranks shard the batch; each rank drives its own library handle (TEST-ONLY host
build) for its samples and their recurrent states.  Checks (1) sharded inference == single-process full batch,
(2) one all-reduce of the flat gradients == gradient of the mean of the per-rank losses."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

NET = {'down_conv_kernels': [[(3, 6)]], 'lstm_kernels': [[(3, 5)]], 'up_conv_kernels': [[(3, 5), (1, 3)]]}
CW = [0.15, 0.25, 0.6]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import lstm_unet_oracle as O
    from tests.emu_backend import emu_session, emu_forward
    from lstm_unet_b200.parallel import shard_range, all_reduce_mean_, max_over_ranks
    GB, T, H, W = 4, 2, 8, 8
    params = O.init_params(NET, seed=3, randomize_bn=True)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((GB, T, 1, H, W)).astype(np.float32)
    lab = rng.integers(-1, 3, size=(GB, T, 1, H, W)).astype(np.float32)
    lo, hi = shard_range(GB, rank, world)
    sess = emu_session(NET, data_format='NCHW', pad_image=False, batch=hi - lo, max_t=T, height=H, width=W,
                       precision='bf16x3', train=True)
    sess.set_params({k: v.numpy().copy() for k, v in params.items()})
    logits, _ = emu_forward(sess, x[lo:hi], False)
    grads = np.zeros(sess.n_trainable, dtype=np.float32)
    loss = np.zeros(1, dtype=np.float32)
    emu_forward(sess, x[lo:hi], True)
    lab_r = np.ascontiguousarray(lab[lo:hi])     # keep alive while the library reads it
    sess.loss_backward(lab_r.ctypes.data, CW, loss.ctypes.data, grads.ctypes.data)
    g = torch.from_numpy(grads)
    all_reduce_mean_(g)
    t = max_over_ranks(1.0 + rank)
    # the same exchange started block by block from inside the backward (parallel.OverlappedAllReduce)
    from lstm_unet_b200.parallel import OverlappedAllReduce
    grads2 = np.zeros(sess.n_trainable, dtype=np.float32)
    g2 = torch.from_numpy(grads2)
    ov = OverlappedAllReduce()
    ov.begin(sess, g2)
    sess.loss_backward(lab_r.ctypes.data, CW, loss.ctypes.data, grads2.ctypes.data)
    ov(g2)
    sess.set_grad_bucket_callback(None)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), logits=logits, grads=g.numpy(), loss=loss, t=t, grads_overlapped=g2.numpy(),
             n_buckets=len(ov.ranges))
    dist.destroy_process_group()


def test_two_rank_sharding_and_gradient_allreduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / ('rank%d.npz' % i)) for i in range(world)]
    assert float(r[0]['t']) == 2.0 and float(r[1]['t']) == 2.0          # max over ranks
    np.testing.assert_array_equal(r[0]['grads'], r[1]['grads'])         # identical after the all-reduce
    np.testing.assert_array_equal(r[0]['grads_overlapped'], r[0]['grads'])   # bucketed exchange == one collective
    assert int(r[0]['n_buckets']) == 2                                  # one Up block + one Down block in this net
    from oracle import lstm_unet_oracle as O
    from tests.emu_backend import emu_session, emu_forward
    params = O.init_params(NET, seed=3, randomize_bn=True)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((4, 2, 1, 8, 8)).astype(np.float32)
    sess = emu_session(NET, data_format='NCHW', pad_image=False, batch=4, max_t=2, height=8, width=8, precision='bf16x3')
    sess.set_params({k: v.numpy().copy() for k, v in params.items()})
    full, _ = emu_forward(sess, x, False)
    np.testing.assert_allclose(np.concatenate([r[0]['logits'], r[1]['logits']], 0), full, rtol=1e-6, atol=1e-7)
    from lstm_unet_b200.parallel import shard_range
    with pytest.raises(ValueError):
        shard_range(5, 0, 2)


NET_BN = {'down_conv_kernels': [[(3, 6), (3, 5)], [(3, 8)]], 'lstm_kernels': [[(3, 5)], [(3, 6)]],
          'up_conv_kernels': [[(3, 7)], [(3, 5), (1, 3)]]}


def _sync_bn_worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import lstm_unet_oracle as O
    from tests.emu_backend import emu_session, emu_forward
    from lstm_unet_b200.parallel import shard_range, all_reduce_mean_, enable_sync_batchnorm
    GB, T, H, W = 4, 2, 8, 8
    params = O.init_params(NET_BN, seed=5, randomize_bn=True)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((GB, T, 1, H, W)).astype(np.float32)
    lab = rng.integers(-1, 3, size=(GB, T, 1, H, W)).astype(np.float32)
    lab[0] = -1                                  # very unequal valid-pixel counts between the ranks
    lo, hi = shard_range(GB, rank, world)
    sess = emu_session(NET_BN, data_format='NCHW', pad_image=False, batch=hi - lo, max_t=T, height=H, width=W,
                       precision='bf16x3', train=True)
    sess.set_params({k: v.numpy().copy() for k, v in params.items()})
    assert enable_sync_batchnorm(sess)
    logits, _ = emu_forward(sess, x[lo:hi], True)
    grads = np.zeros(sess.n_trainable, dtype=np.float32)
    loss = np.zeros(1, dtype=np.float32)
    lab_r = np.ascontiguousarray(lab[lo:hi])
    sess.loss_backward(lab_r.ctypes.data, CW, loss.ctypes.data, grads.ctypes.data)
    g = torch.from_numpy(grads)
    all_reduce_mean_(g)
    moving = {k: v for k, v in sess.get_params().items() if 'moving' in k}
    np.savez(os.path.join(out_dir, 'sync%d.npz' % rank), logits=logits, grads=g.numpy(), loss=loss,
             **{k.replace('/', '|'): v for k, v in moving.items()})
    dist.destroy_process_group()


def test_two_rank_sync_batchnorm_equals_single_device(tmp_path):
    """SURVEY 8e option ii: with the BN statistics summed over the ranks, two ranks holding half the batch each produce
    the logits, moving statistics, loss (mean over the ranks, normalised by the valid pixels of the whole batch) and --
    after the mean all-reduce -- the gradients of ONE device holding the whole batch"""
    world = 2
    mp.spawn(_sync_bn_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / ('sync%d.npz' % i)) for i in range(world)]
    from oracle import lstm_unet_oracle as O
    from tests.emu_backend import emu_session, emu_forward
    params = O.init_params(NET_BN, seed=5, randomize_bn=True)
    rng = np.random.default_rng(1)
    x = rng.standard_normal((4, 2, 1, 8, 8)).astype(np.float32)
    lab = rng.integers(-1, 3, size=(4, 2, 1, 8, 8)).astype(np.float32)
    lab[0] = -1
    sess = emu_session(NET_BN, data_format='NCHW', pad_image=False, batch=4, max_t=2, height=8, width=8, precision='bf16x3',
                       train=True)
    sess.set_params({k: v.numpy().copy() for k, v in params.items()})
    full, _ = emu_forward(sess, x, True)
    grads = np.zeros(sess.n_trainable, dtype=np.float32)
    loss = np.zeros(1, dtype=np.float32)
    sess.loss_backward(lab.ctypes.data, CW, loss.ctypes.data, grads.ctypes.data)
    # (not bit-equal: the shifted partial sums are taken about different pixels and activations are re-split into bf16
    # hi + lo planes, so a 1e-7 difference in the statistics can flip a last bit: the bf16x3 parity tolerance applies)
    np.testing.assert_allclose(np.concatenate([r[0]['logits'], r[1]['logits']], 0), full, rtol=1e-3, atol=1e-4)
    assert abs(0.5 * (float(r[0]['loss'][0]) + float(r[1]['loss'][0])) - float(loss[0])) < 1e-4
    np.testing.assert_array_equal(r[0]['grads'], r[1]['grads'])
    scale = np.abs(grads).max()
    assert np.abs(r[0]['grads'] - grads).max() / scale < 5e-3
    single = sess.get_params()
    for k, v in single.items():
        if 'moving' in k:
            np.testing.assert_allclose(r[0][k.replace('/', '|')], v, rtol=1e-4, atol=1e-6)
            np.testing.assert_array_equal(r[0][k.replace('/', '|')], r[1][k.replace('/', '|')])
