"""GPU diagnostic: run one small network through the scalar mirror engine and the tcgen05 engine and print how the
tensor-core path differs (ConvLSTM states isolate the LSTM kernel; logits cover everything).  One case per process
(a faulting kernel poisons the CUDA context)."""
import sys
import os
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lstm_unet_b200.Networks import ULSTMnet2D
from oracle import lstm_unet_oracle as O

NETS = {
    'one': {'down_conv_kernels': [[(3, 64)]], 'lstm_kernels': [[(5, 64)]], 'up_conv_kernels': [[(3, 32), (1, 3)]]},
    'two': {'down_conv_kernels': [[(3, 64), (3, 64)], [(3, 128), (3, 128)]], 'lstm_kernels': [[(5, 64)], [(5, 128)]],
            'up_conv_kernels': [[(3, 64), (3, 64)], [(3, 32), (3, 32), (1, 3)]]},
    'odd': {'down_conv_kernels': [[(3, 8), (3, 8)], [(3, 12), (3, 12)], [(3, 12), (3, 12)], [(3, 16), (3, 16)]],
            'lstm_kernels': [[(5, 8)], [(5, 12)], [(5, 12)], [(5, 16)]],
            'up_conv_kernels': [[(3, 12), (3, 12)], [(3, 8), (3, 8)], [(3, 8), (3, 8)], [(3, 4), (3, 4), (1, 3)]]},
}


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    net_name, engine, a_mode, precision, T = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
    H, W = (int(sys.argv[6]), int(sys.argv[7])) if len(sys.argv) > 7 else (32, 24)
    net = NETS[net_name]
    B = 2
    params = O.init_params(net, seed=1, randomize_bn=True)
    pnp = {k: v.numpy() for k, v in params.items()}
    x = np.random.default_rng(0).standard_normal((B, T, 1, H, W)).astype(np.float32)
    ora = O.OracleNet(net, 'NCHW', False, params=params)
    ref_l, _ = ora(torch.from_numpy(x), False)
    ref_states = ora.get_states()
    outs = {}
    for eng in (['simt'] if engine == 'simt' else ['simt', engine]):
        m = ULSTMnet2D(net, 'NCHW', False, precision=precision, engine=eng, a_mode=a_mode)
        m.set_weights_dict(pnp)
        logits, _ = m(x, training=False)
        torch.cuda.synchronize()
        outs[eng] = (logits.numpy(), m.get_states())
    tag = '%s/%s/%s/%s/T%d/%dx%d' % (net_name, engine, a_mode, precision, T, H, W)
    for eng, (lg, st) in outs.items():
        print('%s [%s] logits vs oracle: %.3e ; lstm0 h: %.3e c: %.3e'
              % (tag, eng, rel(lg, ref_l.numpy()), rel(st[0][0][0], ref_states[0][0][0]), rel(st[0][0][1], ref_states[0][0][1])))
    if engine != 'simt':
        a, b = outs[engine], outs['simt']
        print('%s tc vs simt: logits %.3e ; lstm0 h %.3e c %.3e' % (tag, rel(a[0], b[0]), rel(a[1][0][0][0], b[1][0][0][0]),
                                                                  rel(a[1][0][0][1], b[1][0][0][1])))
        d = np.abs(a[1][0][0][0] - b[1][0][0][0])          # (B,F,H,W) h of lstm 0
        if d.max() > 1e-3:
            bad = np.argwhere(d > 1e-3)
            print('  bad h entries: %d of %d; first few (b,f,y,x): %s' % (len(bad), d.size, bad[:8].tolist()))
            print('  bad per f (first 16):', (d > 1e-3).sum(axis=(0, 2, 3))[:16].tolist())
            print('  bad per y:', (d > 1e-3).sum(axis=(0, 1, 3)).tolist())
            print('  bad per x:', (d > 1e-3).sum(axis=(0, 1, 2)).tolist())


if __name__ == '__main__':
    main()
