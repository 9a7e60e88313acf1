"""The multi-GPU semantics on the real path (NCCL, tcgen05 kernels): two ranks, one process per GPU.  Self-skipping on a
box with fewer than two devices.  The assertions are those of tests/test_parallel_gloo.py (host build, gloo):
  * batch-sharded inference == one device holding the whole batch;
  * gradient exchange started per block from inside the backward (parallel.OverlappedAllReduce) == one collective;
  * sync_bn=True: BatchNorm statistics and the valid-pixel normaliser of the loss (/root/reference/losses.py:26) over the
    batch of ALL ranks: logits, loss, gradients and moving statistics of one device holding the whole batch."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

NET = {'down_conv_kernels': [[(3, 64), (3, 64)], [(3, 128), (3, 128)]], 'lstm_kernels': [[(5, 64)], [(5, 128)]],
       'up_conv_kernels': [[(3, 64), (3, 64)], [(3, 32), (3, 32), (1, 3)]]}
CW = [0.15, 0.25, 0.6]
GB, T, H, W = 4, 2, 32, 48


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    from oracle import lstm_unet_oracle as O
    params = O.init_params(NET, seed=3, randomize_bn=True)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((GB, T, 1, H, W)).astype(np.float32)
    lab = rng.integers(-1, 3, size=(GB, T, 1, H, W)).astype(np.float32)
    lab[0] = -1                                   # very unequal numbers of annotated pixels per rank
    return {k: v.numpy().copy() for k, v in params.items()}, x, lab


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from lstm_unet_b200.Networks import ULSTMnet2D
    from lstm_unet_b200.parallel import shard_range, all_reduce_mean_, OverlappedAllReduce
    weights, x, lab = _data()
    lo, hi = shard_range(GB, rank, world)
    out = {}
    # (1) sharded inference
    m = ULSTMnet2D(NET, 'NCHW', False, precision='bf16x3')
    m.set_weights_dict(weights)
    out['logits'] = m(x[lo:hi], False)[0].numpy().copy()
    m.close()
    # (2) local-BN data parallel step: overlapped exchange == single collective; 'local' = no exchange at all, run twice
    # (what two identical steps differ by: the order of the fp32 atomics in the BatchNorm statistics and the weight
    # gradients).  On the SMOOTH variant of the network (sigmoid gates, LeakyReLU slope 1): with the reference's kinks a
    # last-bit difference in a BN statistic flips single sub-gradients and two correct runs differ by ~1e-2 in some
    # tensors (tools/grad_sensitivity_probe.py, tools/diag_determinism.py).
    for name, red in (('local_a', None), ('local_b', None), ('single', all_reduce_mean_), ('overlapped', OverlappedAllReduce())):
        m = ULSTMnet2D(NET, 'NCHW', False, precision='bf16x3', train=True, gate='sigmoid', lrelu_alpha=1.0)
        m.set_weights_dict(weights)
        m(x[lo:hi], True)
        if hasattr(red, 'begin'):
            m._grads = torch.zeros(m._sess.n_trainable, dtype=torch.float32, device='cuda')
            red.begin(m._sess, m._grads)
        loss, g = m.backward(lab[lo:hi], CW)
        if name == 'single':
            torch.cuda.synchronize()
            out['grads_single_before'] = g.cpu().numpy().copy()
        if red is not None:
            red(g)
        torch.cuda.synchronize()
        out['grads_' + name] = g.cpu().numpy().copy()
        if hasattr(red, 'ranges'):
            out['n_buckets'] = len(red.ranges)
            out['layout'] = np.array([[e['offset'], e['count']] for e in m._sess.layout if e['trainable']])
            out['names'] = np.array([e['name'] for e in m._sess.layout if e['trainable']])
        m.close()
    # (3) synchronised BatchNorm + global loss normaliser
    m = ULSTMnet2D(NET, 'NCHW', False, precision='bf16x3', train=True, sync_bn=True, gate='sigmoid', lrelu_alpha=1.0)
    m.set_weights_dict(weights)
    lg, _ = m(x[lo:hi], True)
    loss, g = m.backward(lab[lo:hi], CW)
    all_reduce_mean_(g)
    torch.cuda.synchronize()
    out['sync_logits'] = lg.numpy().copy()
    out['sync_loss'] = np.array([float(loss)])
    out['sync_grads'] = g.cpu().numpy().copy()
    for k, v in m.get_weights_dict().items():
        if 'moving' in k:
            out['mv|' + k.replace('/', '|')] = v
    m.close()
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), **out)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two CUDA devices (run under gpurun --gpus 2)')
def test_two_rank_nccl_sharding_overlapped_allreduce_and_sync_bn(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / ('rank%d.npz' % i)) for i in range(world)]
    from lstm_unet_b200.Networks import ULSTMnet2D
    weights, x, lab = _data()
    # single device, whole batch
    m = ULSTMnet2D(NET, 'NCHW', False, precision='bf16x3')
    m.set_weights_dict(weights)
    full = m(x, False)[0].numpy().copy()
    m.close()
    np.testing.assert_allclose(np.concatenate([r[0]['logits'], r[1]['logits']], 0), full, rtol=1e-5, atol=1e-6)
    np.testing.assert_array_equal(r[0]['grads_single'], r[1]['grads_single'])            # identical after the all-reduce
    assert int(r[0]['n_buckets']) == 4                                                    # 2 Up + 2 Down blocks
    def per_tensor(a, b):
        worst = ('', 0.0)
        for (off, cnt), name in zip(r[0]['layout'], r[0]['names']):
            sc = np.abs(b[off:off + cnt]).max()
            if sc < 1e-6:
                continue
            e = float(np.abs(a[off:off + cnt] - b[off:off + cnt]).max() / sc)
            if e > worst[1]:
                worst = (str(name), e)
        return worst
    rerun = max((per_tensor(r[k]['grads_local_a'], r[k]['grads_local_b']) for k in range(world)), key=lambda t: t[1])
    again = max((per_tensor(r[k]['grads_single_before'], r[k]['grads_local_a']) for k in range(world)), key=lambda t: t[1])
    exact = per_tensor(r[0]['grads_single'], 0.5 * (r[0]['grads_single_before'] + r[1]['grads_single_before']))
    print('DIAG rerun per rank', rerun, '| third run vs first', again, '| NCCL AVG vs numpy mean of its inputs', exact)
    mean_local = 0.5 * (r[0]['grads_local_a'] + r[1]['grads_local_a'])
    w_single = per_tensor(r[0]['grads_single'], mean_local)
    w_over = per_tensor(r[0]['grads_overlapped'], mean_local)
    print('two identical local backward passes differ by', rerun, '; single vs mean of locals', w_single, '; overlapped', w_over)
    # same forward, same backward; the weight-gradient atomics add in another order from run to run
    assert rerun[1] < 5e-4, rerun
    assert w_single[1] < 5e-4, w_single
    assert w_over[1] < 5e-4, w_over
    m = ULSTMnet2D(NET, 'NCHW', False, precision='bf16x3', train=True, gate='sigmoid', lrelu_alpha=1.0)
    m.set_weights_dict(weights)
    lg, _ = m(x, True)
    loss, g = m.backward(lab, CW)
    g = g.cpu().numpy()
    np.testing.assert_allclose(np.concatenate([r[0]['sync_logits'], r[1]['sync_logits']], 0), lg.numpy(), rtol=1e-3, atol=1e-4)
    assert abs(0.5 * (float(r[0]['sync_loss'][0]) + float(r[1]['sync_loss'][0])) - float(loss)) < 1e-4
    np.testing.assert_array_equal(r[0]['sync_grads'], r[1]['sync_grads'])
    assert np.abs(r[0]['sync_grads'] - g).max() / np.abs(g).max() < 5e-3
    for k, v in m.get_weights_dict().items():
        if 'moving' in k:
            np.testing.assert_allclose(r[0]['mv|' + k.replace('/', '|')], v, rtol=1e-4, atol=1e-6)
    m.close()
