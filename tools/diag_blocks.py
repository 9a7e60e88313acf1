"""GPU diagnostic: stand-alone UpBlock2D / DownBlock2D against the block oracle over a grid of configurations
(engine, staging mode, precision, sizes, channel counts), one line per case.   python tools/diag_blocks.py"""
import itertools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lstm_unet_b200.Networks import DownBlock2D, UpBlock2D      # noqa: E402
from oracle import blocks_oracle as BO                          # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def up_case(kernels, f, logits, N, h, w, C, Cs, seq, seed=5, rseed=1, **kw):
    ora = BO.OracleUpBlock(kernels, f, 'NCHW', logits, in_channels=C, skip_channels=Cs, seed=seed)
    blk = UpBlock2D(kernels, f, 'NCHW', logits, **kw)
    blk.set_weights_dict({k: v.numpy() for k, v in ora.params.items()})
    rng = np.random.default_rng(rseed)
    errs = []
    for training in seq:
        x = rng.standard_normal((N, C, h, w)).astype(np.float32)
        s = rng.standard_normal((N, Cs, h * f, w * f)).astype(np.float32)
        ref = ora((torch.from_numpy(x), torch.from_numpy(s)), training).numpy()
        got = blk((x, s), training).numpy()
        errs.append('%s %.1e' % ('T' if training else 'F', rel(got, ref)))
    blk.close()
    return ' | '.join(errs)


K3 = [(3, 16), (3, 32), (3, 64)]
if len(sys.argv) > 1 and sys.argv[1] == 'b':
    for seq in ((True, False), (True, True, False), (False, True, False), (True,), (False,)):
        for seed, rseed in ((13, 2), (5, 1), (13, 1), (5, 2)):
            for logits in (False, True):
                print('K3 f2 50x50 seq %-22s seed %2d rseed %d logits %d: %s' % (seq, seed, rseed, logits, up_case(
                    K3, 2, logits, 6, 50, 50, 3, 3, seq, seed=seed, rseed=rseed, precision='bf16x3')), flush=True)
    for nk in (1, 2, 3):
        print('first call training, %d convs: %s' % (nk, up_case(K3[:nk], 1, False, 6, 50, 50, 3, 3, (True, False), seed=13, rseed=2, precision='bf16x3')), flush=True)
        print('first call training, %d convs simt: %s' % (nk, up_case(K3[:nk], 1, False, 6, 50, 50, 3, 3, (True, False), seed=13, rseed=2, precision='bf16x3', engine='simt')), flush=True)
    sys.exit(0)

cases = []
for eng, am, prec in (('tcgen05', 'halo', 'bf16x3'), ('simt', 'halo', 'bf16x3'), ('tcgen05', 'direct', 'bf16x3'), ('tcgen05', 'halo', 'bf16')):
    cases.append(('full F,T,F 50x50 f2', K3, 2, False, 6, 50, 50, 3, 3, (False, True, False), eng, am, prec))
for kern, name in (([(3, 16)], 'conv0 only'), ([(3, 64)], 'conv0 N=64'), ([(1, 16)], 'conv0 1x1')):
    for (C, Cs) in ((3, 3), (64, 3), (3, 64), (5, 3), (3, 5), (64, 64)):
        cases.append(('%s logits C=%d Cs=%d 50x50 f1' % (name, C, Cs), kern, 1, True, 6, 50, 50, C, Cs, (False,), 'tcgen05', 'halo', 'bf16x3'))
cases.append(('conv0 only f2 small 6x10', [(3, 16)], 2, True, 3, 6, 10, 5, 3, (False,), 'tcgen05', 'halo', 'bf16x3'))
cases.append(('conv0 only f1 16x8', [(3, 16)], 1, True, 1, 16, 8, 3, 3, (False,), 'tcgen05', 'halo', 'bf16x3'))
cases.append(('conv0+bn f1 50x50 eval', [(3, 16)], 1, False, 6, 50, 50, 3, 3, (False,), 'tcgen05', 'halo', 'bf16x3'))
cases.append(('conv0+bn f1 50x50 train', [(3, 16)], 1, False, 6, 50, 50, 3, 3, (True, True), 'tcgen05', 'halo', 'bf16x3'))
cases.append(('two convs f1 50x50 eval', [(3, 16), (3, 32)], 1, False, 6, 50, 50, 3, 3, (False,), 'tcgen05', 'halo', 'bf16x3'))
for name, kern, f, lg, N, h, w, C, Cs, seq, eng, am, prec in cases:
    try:
        print('%-44s %-8s %-6s %-7s: %s' % (name, eng, am, prec, up_case(kern, f, lg, N, h, w, C, Cs, seq, engine=eng, a_mode=am, precision=prec)), flush=True)
    except Exception as e:
        print('%-44s %-8s %-6s %-7s: ERROR %s' % (name, eng, am, prec, str(e)[:200]), flush=True)
