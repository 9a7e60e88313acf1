"""How far apart are the gradients of two CORRECT implementations of the reference train step on the CTC network?
CPU only (oracle).  fp32 vs fp64 oracle: ~2e-6.  fp32 oracle whose conv outputs carry 1e-5 relative noise (what a
split-bf16 / different-summation-order forward differs by) vs the fp64 oracle: overall L2 ~1.5e-2, single tensors up to
~1e-1 -- LeakyReLU / hard_sigmoid kinks flip for pre-activations within 1e-5 of them, and at the 8x8 ... 16x16 levels one
flipped element is a per-cent effect on a per-channel sum.  This is why tests/test_gpu_ctc_parity.py compares the CTC-size
backward tightly on the SMOOTH variant of the network (sigmoid gates, LeakyReLU slope 1) and loosely on the reference one.
    python tools/grad_sensitivity_probe.py"""
import os
import sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lstm_unet_oracle as O
torch.set_num_threads(8)
net = O.CTC_NET_PARAMS
CW=[0.15,0.25,0.6]
rng = np.random.default_rng(5)
x = rng.standard_normal((1,2,1,64,64)).astype(np.float32)
lab = rng.integers(-1,3,size=(1,2,1,64,64)).astype(np.float32)
p0 = O.init_params(net, seed=0, randomize_bn=True)
def grads(dtype, noise=0.0, seed=1):
    g = torch.Generator().manual_seed(seed)
    params = {k: v.clone().to(dtype) for k, v in p0.items()}
    ora = O.OracleNet(net, 'NCHW', False, params=params, dtype=dtype)
    if noise:
        orig = O.conv2d_same
        def noisy(x_, w, b, s):
            y = orig(x_, w, b, s)
            return y * (1 + noise * torch.randn(y.shape, generator=g, dtype=y.dtype))
        O.conv2d_same = noisy
    names = ora.trainable_names()
    m = {n: torch.zeros_like(ora.params[n]) for n in names}; v = {n: torch.zeros_like(ora.params[n]) for n in names}
    loss, _, _, gr = O.train_step(ora, torch.from_numpy(x).to(dtype), torch.from_numpy(lab).to(dtype), CW, m, v, 1, 1e-5)
    if noise: O.conv2d_same = orig
    return float(loss), {k: t.double().numpy() for k, t in gr.items()}
l32, g32 = grads(torch.float32)
l64, g64 = grads(torch.float64)
ln, gn = grads(torch.float32, noise=1e-5)
def cmp(a, b, tag):
    num=den=0; worst=('',0)
    for k in a:
        if np.abs(b[k]).max() < 1e-7: continue
        e = np.abs(a[k]-b[k]).max()/np.abs(b[k]).max()
        num += ((a[k]-b[k])**2).sum(); den += (b[k]**2).sum()
        if e > worst[1]: worst=(k,e)
    print(tag, 'worst', worst, 'l2 %.3e' % (num/den)**0.5, flush=True)
cmp(g32, g64, 'fp32 vs fp64')
cmp(gn, g64, 'fp32+1e-5 noise on conv outputs vs fp64')
print('first kernel:', np.abs(gn['DownLayers/0/ConvLSTM/0/kernel']-g64['DownLayers/0/ConvLSTM/0/kernel']).max()/np.abs(g64['DownLayers/0/ConvLSTM/0/kernel']).max())
