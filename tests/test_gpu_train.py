"""Training-step parity on the GPU (tcgen05 forward, data gradients and weight gradients -- lu_wgrad_tc_kernel /
lu_wgrad_pair_kernel) vs autograd through the oracle, the tcgen05 weight gradients vs the scalar engine on identical
activations, and the train2D / Inference2D call mirrors.  Tolerance: 5e-3 relative per gradient tensor in the bf16x3
parity mode (measured ~5e-5); the CTC-size network is in tests/test_gpu_ctc_parity.py."""
import numpy as np
import pytest
import torch

from oracle import lstm_unet_oracle as O
from tests.test_emu_train import NET_B, NET_C, CW, TieWatch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("net,B,T,H,W,seed", [(NET_B, 2, 2, 8, 8, 23), (NET_C, 1, 2, 16, 16, 31)])
def test_train_step_parity(net, B, T, H, W, seed):
    from lstm_unet_b200.Networks import ULSTMnet2D, Adam
    params = O.init_params(net, seed=seed, randomize_bn=True)
    ora = O.OracleNet(net, 'NCHW', False, params=params)
    ora.gate = lambda x: O.hard_sigmoid(x)
    model = ULSTMnet2D(net, 'NCHW', False, precision='bf16x3', train=True)
    model.set_weights_dict({k: v.numpy().copy() for k, v in params.items()})
    opt = Adam(lr=1e-3)
    names = ora.trainable_names()
    m = {n: torch.zeros_like(ora.params[n]) for n in names}
    v = {n: torch.zeros_like(ora.params[n]) for n in names}
    rng = np.random.default_rng(seed)
    compared = 0
    for step in (1, 2):
        x = rng.standard_normal((B, T, 1, H, W)).astype(np.float32)
        lab = rng.integers(-1, 3, size=(B, T, 1, H, W)).astype(np.float32)
        with TieWatch() as tw:
            ref_loss, ref_logits, _, ref_grads = O.train_step(ora, torch.from_numpy(x), torch.from_numpy(lab), CW, m, v, step, 1e-3)
        logits, _ = model(x, True)
        loss, grads = model.backward(lab, CW)
        g = grads.cpu().numpy()
        assert abs(float(loss) - float(ref_loss)) < 1e-4 * max(1.0, abs(float(ref_loss)))
        if tw.clean():
            compared += 1
            for e in model._sess.layout:
                if not e['trainable']:
                    continue
                r = ref_grads[e['name']].numpy()
                mine = g[e['offset']:e['offset'] + e['count']].reshape(e['shape'])
                if np.abs(r).max() < 1e-6:          # conv bias in front of a BatchNorm: analytically zero
                    assert np.abs(mine).max() < 1e-4
                    continue
                err = np.abs(mine - r).max() / np.abs(r).max()
                assert err < 5e-3, (step, e['name'], err)
        model.apply_gradients(grads, opt)
    assert compared == 2


def test_train2d_mirror_runs_and_learns():
    from lstm_unet_b200 import Params, train2D
    net = {'down_conv_kernels': [[(3, 16), (3, 16)], [(3, 32), (3, 32)]], 'lstm_kernels': [[(5, 16)], [(5, 32)]],
           'up_conv_kernels': [[(3, 32), (3, 32)], [(3, 16), (3, 16), (1, 3)]]}
    p = Params.CTCParams({'net_kernel_params': net, 'crop_size': (32, 32), 'batch_size': 2, 'unroll_len': 2,
                          'learning_rate': 1e-3, 'validation_interval': 3, 'print_to_console_interval': 100,
                          'precision': 'bf16x3'})
    train2D.params = p
    losses = train2D.train(num_iterations=8, log=lambda *a: None)
    assert len(losses) == 8 and all(np.isfinite(losses))
    assert min(losses[4:]) < losses[0]              # random labels: the weighted CE still drops from its initial value


def test_inference2d_mirror_streaming_matches_oracle(tmp_path):
    from lstm_unet_b200 import Params, Inference2D
    from lstm_unet_b200.Networks import ULSTMnet2D
    import pickle
    net = NET_B
    params = O.init_params(net, seed=4, randomize_bn=True)
    m0 = ULSTMnet2D(net, 'NCHW', True, precision='bf16x3')
    m0.set_weights_dict({k: v.numpy().copy() for k, v in params.items()})
    m0.save_weights(str(tmp_path / 'model.ckpt'))
    with open(tmp_path / 'model_params.pickle', 'wb') as f:
        pickle.dump({'name': 'ULSTMnet2D', 'params': (net,)}, f)
    frames = [np.random.default_rng(i).standard_normal((20, 28)).astype(np.float32) for i in range(4)]
    p = Params.CTCInferenceParams({'model_path': str(tmp_path), 'pre_sequence_frames': 2, 'precision': 'bf16x3'})
    Inference2D.params = p
    outs = Inference2D.inference(frames)
    assert len(outs) == 4 and outs[0].shape == (3, 20, 28)
    ora = O.OracleNet(net, 'NCHW', True, params=params)
    seq = frames[:2][::-1] + frames
    ref = [ora(torch.from_numpy(f).reshape(1, 1, 1, 20, 28), False)[1][0, 0].numpy() for f in seq][2:]
    for a, b in zip(outs, ref):
        assert np.abs(a - b).max() < 1e-3


@pytest.mark.parametrize("precision,tol", [('bf16x3', 5e-5), ('bf16', 5e-5)])
def test_wgrad_tcgen05_matches_scalar_wgrad_wide(precision, tol):
    """Weight gradients of a network with full 64-channel chunks (pairs of activation stages, 128-column slabs,
    stride-2 parity planes, two-source convs, patches): the tcgen05 weight-gradient kernel vs the scalar engine on the
    SAME forward activations and upstream gradients (LU_WGRAD_ENGINE switches only the weight-gradient engine; the
    backward is idempotent), so the comparison is free of activation-kink noise.  Both accumulate the same bf16
    products in fp32; only the summation order differs."""
    import os
    from lstm_unet_b200.Networks import ULSTMnet2D
    net = {'down_conv_kernels': [[(3, 64), (3, 64)], [(3, 128), (3, 128)], [(3, 192), (3, 192)]],
           'lstm_kernels': [[(5, 64)], [(5, 128)], [(3, 192)]],
           'up_conv_kernels': [[(3, 128), (3, 128)], [(3, 64), (3, 64)], [(3, 32), (3, 32), (1, 3)]]}
    params = O.init_params(net, seed=2, randomize_bn=True)
    rng = np.random.default_rng(0)
    B, T, H, W = 2, 3, 48, 40
    m = ULSTMnet2D(net, 'NCHW', False, precision=precision, train=True)
    m.set_weights_dict({k: v.numpy().copy() for k, v in params.items()})
    for call in range(2):                       # second call: non-zero initial states (h_init pass of the recurrent wgrad)
        x = rng.standard_normal((B, T, 1, H, W)).astype(np.float32)
        lab = rng.integers(-1, 3, size=(B, T, 1, H, W)).astype(np.float32)
        m(x, True)
        os.environ.pop('LU_WGRAD_ENGINE', None)
        l1, g = m.backward(lab, CW)
        g1 = g.cpu().numpy().copy()
        os.environ['LU_WGRAD_ENGINE'] = 'simt'
        try:
            l0, g = m.backward(lab, CW)
            g0 = g.cpu().numpy().copy()
        finally:
            os.environ.pop('LU_WGRAD_ENGINE', None)
        assert float(l0) == float(l1)
        for e in m._sess.layout:
            if not e['trainable']:
                continue
            a, b = g0[e['offset']:e['offset'] + e['count']], g1[e['offset']:e['offset'] + e['count']]
            scale = np.abs(a).max()
            if scale < 1e-6 or (e['name'].endswith('bias') and scale < 1e-4):
                continue          # conv biases in front of a BatchNorm: analytically zero, atomics-order noise
            assert np.abs(a - b).max() / scale < tol, (call, e['name'], np.abs(a - b).max() / scale)


def test_train2d_mirror_saves_like_the_reference(tmp_path):
    """train2D.py:222-240: a checkpoint at the last step, a checkpoint when the loop dies of ValueError / KeyboardInterrupt,
    and the inference model (model.ckpt + model_params.pickle) in a `finally`, whatever happened."""
    import os
    from lstm_unet_b200 import Params, train2D, checkpoint as ck
    net = {'down_conv_kernels': [[(3, 16)], [(3, 32)]], 'lstm_kernels': [[(3, 16)], [(3, 32)]],
           'up_conv_kernels': [[(3, 16)], [(3, 16), (1, 3)]]}
    base = {'net_kernel_params': net, 'crop_size': (32, 32), 'batch_size': 2, 'unroll_len': 2, 'learning_rate': 1e-3,
            'validation_interval': 100, 'print_to_console_interval': 100, 'save_checkpoint_iteration': 4, 'dry_run': False,
            'save_checkpoint_dir': str(tmp_path), 'save_log_dir': str(tmp_path), 'seed': 3}
    p = Params.CTCParams(dict(base, experiment_name='full'))
    train2D.params = p
    train2D.train(num_iterations=6, log=lambda *a: None)
    d = p.experiment_save_dir
    assert os.path.exists(os.path.join(d, 'model.ckpt.index')) and os.path.exists(os.path.join(d, 'model_params.pickle'))
    assert ck.latest_checkpoint(os.path.join(d, 'tf_ckpts')).endswith('ckpt-6')          # step 4, and the final step 6

    p = Params.CTCParams(dict(base, experiment_name='interrupted'))
    train2D.params = p
    prov = p.train_data_provider
    real, calls = prov.get_batch, []

    def flaky():
        calls.append(1)
        if len(calls) == 4:
            raise ValueError('queue closed')
        return real()
    prov.get_batch = flaky
    losses = train2D.train(num_iterations=10, log=lambda *a: None)
    assert len(losses) == 3
    d = p.experiment_save_dir
    assert ck.latest_checkpoint(os.path.join(d, 'tf_ckpts')).endswith('ckpt-3')
    assert os.path.exists(os.path.join(d, 'model.ckpt.index'))
    # continue_run: the next run writes into the directory of the run it continues (Params.py:132-142)
    p2 = Params.CTCParams(dict(base, experiment_name='interrupted', load_checkpoint=True, continue_run=True,
                               load_checkpoint_path=os.path.join(d, 'tf_ckpts')))
    assert p2.experiment_save_dir in (d, os.path.join(d, 'tf_ckpts'))
    train2D.params = p2
    more = train2D.train(num_iterations=5, log=lambda *a: None)
    assert len(more) == 2                                                                  # resumed at step 3
