#!/usr/bin/env python
"""Benchmark of the ConvLSTM-UNet hot path: frames/sec at 512x512, T=8, batch 4 per GPU (BASELINE.json configs[1]:
inference-only forward through the Inference2D model call: pad_image=True, training=False, stateful h/c carried
between iterations).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode infer]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one model call over a (4, 8, 1, 512, 512) batch of synthetic z-scored frames (N(0,1), DataHandeling.py:103).
Prints ONE JSON line (rank 0).  `value` = frames/s with the batch resident in HBM; `e2e` = the same through the public
API with HOST buffers (pinned H2D of the frames + D2H of the soft-max, inside the timed region).  `roofline` is for
the dominant kernel (the tcgen05 ConvLSTM step kernel, 93 % of the FLOPs), timed live with CUDA events around every
launch.  `cpu_baseline` times the CPU restatement of the reference (oracle/, torch-CPU: TensorFlow, which the reference
needs, is not installable in this image) on a bounded sample.  `--impl reference` runs only that CPU arm.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'frames/sec (512x512, T=8, bf16)'
CTC_NET = {   # Params.py:49-69
    'down_conv_kernels': [[(3, 128), (3, 128)], [(3, 256), (3, 256)], [(3, 256), (3, 256)], [(3, 512), (3, 512)]],
    'lstm_kernels': [[(5, 128)], [(5, 256)], [(5, 256)], [(5, 512)]],
    'up_conv_kernels': [[(3, 256), (3, 256)], [(3, 128), (3, 128)], [(3, 64), (3, 64)], [(3, 32), (3, 32), (1, 3)]],
}


def profile_traffic_bytes(name='r2_ncu_prof_lstm_l1_pair.txt'):
    """DRAM bytes (read + written) of the launch captured in a committed `ncu --set full` summary (profiles/), or None."""
    p = os.path.join(ROOT, 'profiles', name)
    try:
        tot = 0.0
        for line in open(p):
            f = line.split()
            if f and f[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                tot += float(f[2]) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}[f[1]]
        return tot or None
    except Exception:
        return None


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('bf16_tflops_sustained', d.get('bf16_tflops', 1590.0)), d.get('bf16_tflops', 1590.0), 'measured'
    return 1400.0, 1590.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        """(Re)start sampling: called around every timed region, so that the reported median is the clock UNDER LOAD."""
        q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, args=(self.proc,), daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def pause(self):
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
            self.thread.join(timeout=2)
            self.proc = None
            self.seen = True

    def _read(self, proc):
        for line in proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        self.pause()
        if not getattr(self, 'seen', False):
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for j, nme in enumerate(names):
                    if r[5 + j].lower().startswith('active'):
                        reasons.add(nme)
            except Exception:
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def workload_config(args, world, training=False):
    """`config` of the JSON line: names the workload only, identical for `--impl ours` and `--impl reference`."""
    B, T, S = args.batch, args.unroll, args.size
    c5 = (S, T, B) == (1024, 16, 1)                  # BASELINE.json configs[4]: long-sequence stress, one sequence per GPU
    if args.mode == 'train' or training:
        wl = (('C5' if c5 else 'C3/C4') + ': ConvLSTM-UNet (CTCParams net, 74.6M params) full train step (fwd + WeightedCELoss + bwd + gradient '
              'all-reduce + Adam), %dx%d, T=%d, batch %d per GPU, pad_image=False, stateful' % (S, S, T, B))
    else:
        wl = (('C5 (inference)' if c5 else 'C2') + ': ConvLSTM-UNet (CTCParams net, 74.6M params) inference forward, %dx%d, T=%d, batch %d per GPU, '
              'pad_image=True, stateful' % (S, S, T, B))
    return {'workload': wl, 'global_batch': B * world,
            'l2_policy': 'inputs+activations per step (>6 GB) exceed the 126 MB L2; no flush needed'}


def _oracle():
    import torch
    from oracle import lstm_unet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return torch, O, cores


def cpu_infer_run(steps, warmup, B=1, T=2, H=512, W=512):
    """CPU restatement of the reference forward (Inference2D call: pad_image=True, training=False, stateful) on a
    bounded sample of the C2 workload: the same network and the same 512x512 frames (528x528 after pad_image), fewer
    of them per step (B=1, T=2 instead of B=4, T=8).  frames/s needs no rescaling."""
    torch, O, cores = _oracle()
    net = O.OracleNet(O.CTC_NET_PARAMS, 'NCHW', True, seed=0)
    x = torch.randn(B, T, 1, H, W)
    with torch.no_grad():
        for _ in range(warmup):
            net(x, False)
        t0 = time.perf_counter()
        for _ in range(steps):
            net(x, False)
        dt = time.perf_counter() - t0
    return {'value': B * T * steps / dt, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'sample': 'oracle (torch-CPU fp32 restatement of the reference; TensorFlow is not installable) inference forward, '
                      'CTC net, %dx%d frames, pad_image, stateful, B=%d T=%d per step (C2 has B=4 T=8), %d steps after %d '
                      'warm-up; no rescaling' % (H, W, B, T, steps, warmup),
            'ms_per_step': dt / steps * 1e3}


def cpu_full_c2_step(B=4, T=8, H=512, W=512):
    """ONE inference step at the true C2 shape on the host cores (about a minute): shows the bounded sample's frames/s
    is the full shape's."""
    torch, O, cores = _oracle()
    net = O.OracleNet(O.CTC_NET_PARAMS, 'NCHW', True, seed=0)
    x = torch.randn(B, T, 1, H, W)
    with torch.no_grad():
        t0 = time.perf_counter()
        net(x, False)
        dt = time.perf_counter() - t0
    return {'value': B * T / dt, 'unit': 'frames/s', 'ms_per_step': dt * 1e3, 'cores': cores,
            'shape': 'B=%d T=%d %dx%d pad_image (true C2 shape), 1 step, no warm-up' % (B, T, H, W)}


def cpu_train_run(steps=1, warmup=1, B=2, T=4, H=128, W=128):
    """BASELINE.json configs[0] on the host cores: train2D.py's train step (forward(training=True) + WeightedCELoss +
    backward + Keras Adam, train2D.py:87-93) at 128x128, T=4, batch 2, through the oracle."""
    torch, O, cores = _oracle()
    net = O.OracleNet(O.CTC_NET_PARAMS, 'NCHW', False, seed=0)
    names = net.trainable_names()
    m = {n: torch.zeros_like(net.params[n]) for n in names}
    v = {n: torch.zeros_like(net.params[n]) for n in names}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, T, 1, H, W, generator=g)
    lab = torch.randint(-1, 3, (B, T, 1, H, W), generator=g).to(torch.float32)
    for i in range(warmup):
        O.train_step(net, x, lab, [0.15, 0.25, 0.6], m, v, i + 1, 1e-5)
    t0 = time.perf_counter()
    for i in range(steps):
        O.train_step(net, x, lab, [0.15, 0.25, 0.6], m, v, warmup + i + 1, 1e-5)
    dt = time.perf_counter() - t0
    return {'value': B * T * steps / dt, 'unit': 'frames/s', 'ms_per_step': dt / steps * 1e3, 'cores': cores, 'kind': 'port',
            'sample': 'C1 (BASELINE.json configs[0]): oracle train step fwd+loss+bwd+Adam, CTC net, %dx%d, T=%d, batch %d, '
                      '%d step(s) after %d warm-up' % (H, W, T, B, steps, warmup)}


def run_reference(args):
    """The reference arm: the CPU restatement of the reference (oracle/; the reference itself needs TensorFlow) with all
    host threads, exactly --steps timed steps after --warmup untimed ones, each step a bounded sample of the workload
    (same network, same frame size, fewer frames); plus ONE step at the true C2 shape and the C1 train step."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    training = args.mode == 'train'
    if training:
        cb = cpu_train_run(max(1, args.steps), max(0, args.warmup), B=1, T=2, H=args.size, W=args.size)
        cb['sample'] = ('oracle train step fwd+loss+bwd+Adam, CTC net, %dx%d frames, B=1 T=2 per step (C3 has B=4 T=8), %d steps '
                        'after %d warm-up' % (args.size, args.size, args.steps, args.warmup))
    else:
        cb = cpu_infer_run(max(1, args.steps), max(0, args.warmup), H=args.size, W=args.size)
    line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': cb['ms_per_step'], 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, max(1, args.gpus)),
            'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': cb['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    if not training and not args.no_full:
        line['full_shape'] = cpu_full_c2_step(args.batch, args.unroll, args.size, args.size)
        tr = cpu_train_run()
        line['train'] = {k: tr[k] for k in ('value', 'unit', 'ms_per_step', 'cores', 'kind', 'sample')}
    emit(line)


CW = [0.15, 0.25, 0.6]          # Params.py:72 class_weights


def parity_spot_check(rank):
    """Pre-timing check of the timed path against the oracle with the SAME weights the timed model uses (Keras default
    initialisation, seed 0): CTC network, one stateful sequence B=1, T=4 at 128x128 with pad_image, every operand
    precision; and one train step (loss + all gradients) at B=1, T=2, 64x64.  max-abs error over max-abs reference."""
    if rank != 0:
        return None
    import torch
    from oracle import lstm_unet_oracle as O
    from lstm_unet_b200.Networks import ULSTMnet2D, keras_default_init
    from lstm_unet_b200 import _lib
    torch.set_num_threads(os.cpu_count() or 1)

    def rel(a, b):
        return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / max(np.abs(b).max(), 1e-30))
    out = {'network': 'CTCParams.net_kernel_params, Keras default init seed 0 (the timed models\' weights)',
           'metric': 'max|ours - oracle| / max|oracle|', 'north_star_tolerance': 1e-3}
    rng = np.random.default_rng(11)
    x = rng.standard_normal((1, 4, 1, 128, 128)).astype(np.float32)
    weights = None
    ref = None
    inf = {}
    for prec in [p for p in ('bf16', 'fp16', 'bf16x3') if p in _lib.PRECISIONS]:
        m = ULSTMnet2D(CTC_NET, 'NCHW', True, precision=prec, seed=0, cuda_graph=False)
        lg, sm = m(x, False)
        if weights is None:
            weights = m.get_weights_dict()
            ora = O.OracleNet(O.CTC_NET_PARAMS, 'NCHW', True, params={k: torch.from_numpy(v.copy()) for k, v in weights.items()})
            with torch.no_grad():
                ref = [t.numpy() for t in ora(torch.from_numpy(x), False)]
        inf[prec] = {'logits': rel(lg.numpy(), ref[0]), 'softmax': rel(sm.numpy(), ref[1])}
        m.close()
    out['inference'] = {'shape': 'B=1 T=4 128x128 pad_image', **inf}
    # train step: reference configuration (bf16 = the timed dtype, bf16x3), and the smooth variant of the network
    # (sigmoid gates, LeakyReLU slope 1) where no sub-gradient can flip for a 1e-5 forward difference
    xt = rng.standard_normal((1, 2, 1, 64, 64)).astype(np.float32)
    lab = rng.integers(-1, 3, size=(1, 2, 1, 64, 64)).astype(np.float32)
    tr = {}
    for tag, prec, smooth in (('bf16', 'bf16', False), ('bf16x3', 'bf16x3', False), ('bf16x3_smooth_network', 'bf16x3', True)):
        gate = 'sigmoid' if smooth else 'hard_sigmoid'
        ora = O.OracleNet(O.CTC_NET_PARAMS, 'NCHW', False, gate=gate,
                          params={k: torch.from_numpy(v.copy()) for k, v in weights.items()})
        names = ora.trainable_names()
        mm = {n: torch.zeros_like(ora.params[n]) for n in names}
        vv = {n: torch.zeros_like(ora.params[n]) for n in names}
        saved = O.LRELU_ALPHA
        O.LRELU_ALPHA = 1.0 if smooth else saved
        try:
            ref_loss, _, _, ref_grads = O.train_step(ora, torch.from_numpy(xt), torch.from_numpy(lab), CW, mm, vv, 1, 1e-5)
        finally:
            O.LRELU_ALPHA = saved
        m = ULSTMnet2D(CTC_NET, 'NCHW', False, precision=prec, seed=0, train=True, gate=gate, lrelu_alpha=1.0 if smooth else 0.3)
        m.set_weights_dict(weights)
        m(xt, True)
        loss, grads = m.backward(lab, CW)
        g = grads.cpu().numpy()
        worst, worst_name, num, den = 0.0, '', 0.0, 0.0
        for e in m._sess.layout:
            if not e['trainable']:
                continue
            if '/Conv/' in e['name'] and e['name'].endswith('bias') and not e['name'].startswith('UpLayers/3/Conv/2'):
                continue                     # conv bias in front of a training-mode BatchNorm: analytically zero
            r = ref_grads[e['name']].numpy().reshape(-1)
            mine = g[e['offset']:e['offset'] + e['count']]
            err = rel(mine, r)
            num += float(((mine - r) ** 2).sum()); den += float((r ** 2).sum())
            if err > worst:
                worst, worst_name = err, e['name']
        tr[tag] = {'loss': abs(float(loss) - float(ref_loss)) / abs(float(ref_loss)), 'worst_gradient_tensor': worst,
                   'worst_gradient_name': worst_name, 'all_gradients_l2': (num / den) ** 0.5}
        m.close()
    tr['note'] = ('reference configuration: gradients of two correct implementations whose forwards differ by 1e-5 differ by '
                  '~1e-2 (LeakyReLU / hard_sigmoid sub-gradients flip; tools/grad_sensitivity_probe.py), so the tight check '
                  'is the smooth variant')
    out['train_step'] = {'shape': 'B=1 T=2 64x64', **tr}
    torch.cuda.empty_cache()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')      # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import __graft_entry__ as ge
    if world > 1:                      # one rank (re)builds the in-tree library if it is stale, the others wait
        if rank == 0:
            ge.build()
        dist.barrier()
    ge.build()
    from lstm_unet_b200.Networks import ULSTMnet2D

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(list(vals), device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    B, T, H, W = args.batch, args.unroll, args.size, args.size
    sustained, burst, which = peaks()
    rng = np.random.default_rng(1234 + rank)
    x_host = rng.standard_normal((B, T, 1, H, W)).astype(np.float32)
    x_dev = torch.from_numpy(x_host).cuda()
    parity = parity_spot_check(rank) if not args.no_parity else None
    barrier()
    sampler = ClockSampler(local)        # rank 0 samples nvidia-smi during the timed regions only

    # ------------------------------------------------------------------ C2: inference forward (Inference2D call)
    def time_infer(precision, with_e2e):
        model = ULSTMnet2D(CTC_NET, 'NCHW', pad_image=True, precision=precision, a_mode=args.a_mode, seed=0,
                           cuda_graph={'auto': 'auto', 'on': True, 'off': False}[args.cuda_graph])
        model(x_dev, False)
        torch.cuda.synchronize()
        sess = model._sess
        if rank == 0 and with_e2e:
            sampler.start()              # from the warm-up on (the same load): nvidia-smi needs a moment to start
        for _ in range(args.warmup):
            model(x_dev, False)
        barrier()
        sess.launch_count(reset=True)
        sess.kernel_times(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            model(x_dev, False)          # states carry over between iterations, like the inference loop
        ev1.record()
        barrier()
        if rank == 0 and with_e2e:
            sampler.pause()
        ms = ev0.elapsed_time(ev1)
        launches = sess.launch_count()
        kt = sess.kernel_times(False)
        res = {'ms': ms, 'launches': launches, 'kt': kt, 'flops_step': sess.forward_flops(T) * B,
               'lstm_flops_step': sess.lstm_flops(T) * B, 'graph': bool(model.graph_active)}
        if with_e2e:
            # end to end through the public API with HOST buffers: every step's frames go host -> device from pinned
            # memory and its soft-max comes back device -> host (Inference2D.py:59-60), both inside the timed region.
            # (a) the reference's own per-call form: model(x_host) then .numpy()
            post = None
            if args.post:
                from lstm_unet_b200.postprocess import PostProcessor
                post = PostProcessor()

            def once():
                o = model(x_host, False)
                return post(o[1]).numpy() if post is not None else o[1].numpy()
            for _ in range(3):
                sm = once()
            barrier()
            n_e2e = max(2, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                sm = once()
            barrier()
            serial_ms = (time.perf_counter() - t0) * 1e3
            # (b) the pipelined form of the same public API: copies of batch i+1 / i-1 overlap the compute of batch i
            for _ in model.predict_batches(x_host for _ in range(3)):
                pass
            barrier()
            t0 = time.perf_counter()
            for sm2 in model.predict_batches(x_host for _ in range(n_e2e)):
                pass
            barrier()
            pipe_ms = (time.perf_counter() - t0) * 1e3
            res.update({'e2e_serial_ms': serial_ms, 'e2e_pipe_ms': pipe_ms, 'e2e_steps': n_e2e, 'd2h': int(sm.nbytes),
                        'labelled': post is not None})
        model.close()
        del model
        torch.cuda.empty_cache()
        return res

    # ------------------------------------------------------------------ C3 / C4: full train step
    def time_train():
        from lstm_unet_b200.Networks import Adam
        from lstm_unet_b200.parallel import all_reduce_mean_, OverlappedAllReduce
        model = ULSTMnet2D(CTC_NET, 'NCHW', pad_image=False, precision='bf16', a_mode=args.a_mode, seed=0, train=True,
                           sync_bn=bool(args.sync_bn))
        lab_host = rng.integers(-1, 3, size=(B, T, 1, H, W)).astype(np.float32)
        lab_dev = torch.from_numpy(lab_host).cuda()
        opt = Adam(lr=1e-5)
        reducers = {'none': None}
        if world > 1:
            reducers = {'overlapped': OverlappedAllReduce(), 'single': all_reduce_mean_, 'none': None}
        main = args.allreduce if world > 1 else 'none'

        def step(red, x=x_dev, lab=lab_dev):
            return model.train_step(x, lab, CW, opt, reducers[red])
        step(main)
        torch.cuda.synchronize()
        sess = model._sess
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            step(main)
        barrier()
        sess.launch_count(reset=True)
        sess.kernel_times(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            step(main)
        ev1.record()
        barrier()
        if rank == 0:
            sampler.pause()
        ms = ev0.elapsed_time(ev1)
        launches = sess.launch_count()
        kt = sess.kernel_times(False)
        # exposed cost of the gradient exchange: the same step with no exchange / one collective after the backward
        variants = {}
        if world > 1:
            n_var = max(2, min(args.steps, 5))
            for name in ('none', 'single', 'overlapped'):
                step(name)
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(n_var):
                    step(name)
                b.record()
                barrier()
                variants[name] = a.elapsed_time(b) / n_var
        # e2e: frames and labels from pinned host memory every step, loss read back (train2D.py:103 feeds the metrics)
        for _ in range(2):
            float(step(main, x_host, lab_host)[2])
        barrier()
        n_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            loss = float(step(main, x_host, lab_host)[2])
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        res = {'ms': ms, 'launches': launches, 'kt': kt, 'class_flops': sess.class_flops(T), 'variants': variants,
               'e2e_ms': e2e_ms, 'e2e_steps': n_e2e, 'loss': loss, 'grad_bytes': int(sess.n_trainable) * 4,
               'flops_step': sess.forward_flops(T) * B * 3, 'main': main}
        model.close()
        del model
        torch.cuda.empty_cache()
        return res

    def roofline_of(cls_name, kernel, kt, flops_step, steps, ms_total, extra=None):
        k_ms, k_n = kt[cls_name]
        tf = flops_step * steps / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None
        r = {'bound': 'tensor', 'achieved': tf, 'peak': sustained, 'unit': 'TFLOP/s', 'frac': (tf / sustained) if tf else None,
             'traffic': None, 'kernel': '%s (%d launches)' % (kernel, k_n), 'kernel_ms_per_step': k_ms / steps,
             'kernel_share_of_step': k_ms / ms_total if ms_total else None,
             'peak_source': which + ' bf16_tflops_sustained (kernel timed inside a long step); burst %.1f' % burst}
        if extra:
            r.update(extra)
        return r

    do_infer = args.mode in ('all', 'infer', 'stream')
    do_train = args.mode in ('all', 'train')
    inf = time_infer(args.precision, True) if do_infer else None
    variants = {}
    if do_infer and args.mode == 'all' and not args.no_variants:
        for prec in ('fp16', 'bf16x3'):
            if prec != args.precision:
                variants[prec] = time_infer(prec, False)
    trn = time_train() if do_train else None
    clocks = sampler.stop() if rank == 0 else None

    vals = []
    if inf:
        vals += [inf['ms'], inf['e2e_serial_ms'], inf['e2e_pipe_ms']] + [variants[k]['ms'] for k in sorted(variants)]
    if trn:
        vals += [trn['ms'], trn['e2e_ms']] + [trn['variants'][k] for k in sorted(trn['variants'])]
    red = max_ranks(vals)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    it = iter(red)
    line = {}
    if inf:
        ms, serial_ms, pipe_ms = next(it), next(it), next(it)
        var_ms = {k: next(it) for k in sorted(variants)}
        frames = B * T * world
        tf = roofline_of('lstm_fwd', 'lu_conv_tc_kernel<LSTM> forward launches, all 4 ConvLSTM levels', inf['kt'],
                         inf['lstm_flops_step'], args.steps, ms,
                         {'traffic': profile_traffic_bytes(),
                          'traffic_note': 'DRAM bytes of one level-1 ConvLSTM launch (ncu --set full, profiles/r2_ncu_prof_lstm_l1_pair.txt; '
                                          'algorithmic: x 0.14 + h_in 0.14 + h_out 0.14 + c read/write 0.57 + weights 0.02 = 1.0 GB)',
                          'whole_step_tflops': inf['flops_step'] * args.steps / (ms * 1e-3) / 1e12})
        cfg = workload_config(args, world)
        cfg.update({'parallelism': 'batch-sharded replicas x%d (no data-path collective)' % world, 'a_mode': args.a_mode,
                    'step_tflop': inf['flops_step'] / 1e12, 'cuda_graph': inf['graph'],
                    'switches': {k: v for k, v in sorted(os.environ.items()) if k.startswith('LU_')}})
        dt = {'bf16': 'bf16', 'fp16': 'fp16', 'bf16x3': 'bf16x3(split-bf16, fp32-equivalent)'}
        line = {
            'metric': METRIC, 'value': frames * args.steps / (ms * 1e-3), 'unit': 'frames/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': dt[args.precision], 'data': 'synthetic', 'config': cfg,
            'e2e': {'value': frames * inf['e2e_steps'] / (pipe_ms * 1e-3), 'unit': 'frames/s',
                    'h2d_bytes_per_step': int(x_host.nbytes), 'd2h_bytes_per_step': inf['d2h'], 'steps': inf['e2e_steps'],
                    'ms_per_step': pipe_ms / inf['e2e_steps'],
                    'api': 'ULSTMnet2D.predict_batches(host batches) -> host soft-max per batch (copies of batch i+1 / i-1 '
                           'overlap the compute of batch i; every byte still crosses PCIe inside the timed region)',
                    'per_call_form': {'value': frames * inf['e2e_steps'] / (serial_ms * 1e-3), 'unit': 'frames/s',
                                      'ms_per_step': serial_ms / inf['e2e_steps'],
                                      'api': 'model(x_host, training=False)[1].numpy() per step (Inference2D.py:59-60), copies serial with the compute'},
                    'labelled_on_device': inf['labelled']},
            'gpu_launches': int(inf['launches']), 'clocks': clocks, 'roofline': tf,
        }
        if variants:
            line['dtype_variants'] = {k: {'value': frames * args.steps / (var_ms[k] * 1e-3), 'unit': 'frames/s',
                                          'ms_per_step': var_ms[k] / args.steps, 'dtype': dt[k]} for k in variants}
            line['dtype_variants'][args.precision] = {'value': line['value'], 'unit': 'frames/s',
                                                      'ms_per_step': line['ms_per_step'], 'dtype': dt[args.precision]}
    if trn:
        ms, e2e_ms = next(it), next(it)
        var_ms = {k: next(it) for k in sorted(trn['variants'])}
        frames = B * T * world
        cf = trn['class_flops']
        wg_traffic = {'traffic': profile_traffic_bytes('r2_ncu_prof_wgrad_pair.txt'),
                      'traffic_note': 'DRAM bytes of the level-0 ConvLSTM weight-gradient launch (frames t >= 1 of the recurrent term: '
                                      '28 of 32 frames; profiles/r2_ncu_prof_wgrad_pair.txt; algorithmic: h 1.9 + dz 7.5 GB read once, '
                                      '6.6 MB of fp32 gradient atomics)'}
        blk = {
            'workload': workload_config(args, world, training=True)['workload'],
            'value': frames * args.steps / (ms * 1e-3), 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'dtype': 'bf16', 'step_tflop': trn['flops_step'] / 1e12,
            'whole_step_tflops': trn['flops_step'] * args.steps / (ms * 1e-3) / 1e12,
            'whole_step_frac_of_sustained': trn['flops_step'] * args.steps / (ms * 1e-3) / 1e12 / sustained,
            'sync_bn': bool(args.sync_bn), 'final_loss': trn['loss'], 'gpu_launches': int(trn['launches']),
            'parallelism': ('data parallel x%d: batch-sharded, NCCL mean all-reduce of the %.0f MB fp32 gradient buffer per step (%s)'
                            % (world, trn['grad_bytes'] / 1e6, trn['main'])) if world > 1 else 'single GPU (no exchange)',
            'allreduce': {'bytes_per_step': trn['grad_bytes'] if world > 1 else 0, 'mode': trn['main'],
                          'ms_per_step_by_mode': var_ms,
                          'allreduce_exposed_ms': {k: var_ms[k] - var_ms['none'] for k in var_ms if k != 'none'} if var_ms else None},
            'e2e': {'value': frames * trn['e2e_steps'] / (e2e_ms * 1e-3), 'unit': 'frames/s',
                    'h2d_bytes_per_step': int(x_host.nbytes) * 2, 'd2h_bytes_per_step': 4, 'steps': trn['e2e_steps'],
                    'ms_per_step': e2e_ms / trn['e2e_steps'],
                    'api': 'ULSTMnet2D.train_step(host frames, host labels, ...) + float(loss) (train2D.py:87-103)'},
            'roofline': roofline_of('wgrad', 'lu_wgrad_pair_kernel + lu_wgrad_tc_kernel (weight gradient: dominant kernel class of the train step)',
                                    trn['kt'], cf['wgrad'], args.steps, ms, wg_traffic),
            'rooflines': {
                'wgrad': roofline_of('wgrad', 'lu_wgrad_pair_kernel + lu_wgrad_tc_kernel', trn['kt'], cf['wgrad'], args.steps, ms, wg_traffic),
                'dgrad': roofline_of('dgrad', 'lu_conv_tc_kernel<GRAD>', trn['kt'], cf['dgrad'], args.steps, ms,
                                     {'traffic': profile_traffic_bytes('r2_ncu_prof_dgrad_pair.txt'),
                                      'traffic_note': 'DRAM bytes of one level-1 recurrent data-gradient launch (one time step; '
                                                      'profiles/r2_ncu_prof_dgrad_pair.txt; algorithmic: dz 0.54 + dh 0.13 r/w + weights 0.01 GB)'}),
                'lstm_fwd': roofline_of('lstm_fwd', 'lu_conv_tc_kernel<LSTM>', trn['kt'], cf['lstm_fwd'], args.steps, ms),
                'conv_fwd': roofline_of('conv_fwd', 'lu_conv_tc_kernel<CONV>', trn['kt'], cf['conv_fwd'], args.steps, ms),
            },
        }
        tc_ms = sum(trn['kt'][k][0] for k in trn['kt']) / args.steps
        blk['elementwise_and_other_ms_per_step'] = ms / args.steps - tc_ms
        if inf:
            line['train'] = blk
        else:
            cfg = workload_config(args, world, training=True)
            cfg.update({'parallelism': blk['parallelism'], 'a_mode': args.a_mode, 'step_tflop': blk['step_tflop'],
                        'switches': {k: v for k, v in sorted(os.environ.items()) if k.startswith('LU_')}})
            line = {'metric': METRIC, 'value': blk['value'], 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
                    'warmup': args.warmup, 'ms_per_step': blk['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
                    'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic', 'config': cfg, 'e2e': blk['e2e'],
                    'gpu_launches': blk['gpu_launches'], 'clocks': clocks, 'roofline': blk['roofline'], 'train': blk}
    if parity is not None:
        line['parity_err'] = parity
    if world == 1 and not args.no_cpu:
        cb = cpu_infer_run(2, 1, H=H, W=W) if inf else None       # ~15 s of host work
        ct = cpu_train_run()                                        # ~20 s
        if cb is not None:
            line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
            line['cpu_baseline']['train'] = {k: ct[k] for k in ('value', 'unit', 'ms_per_step', 'sample')}
        else:
            line['cpu_baseline'] = {k: ct[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_postprocess(args):
    """--mode postprocess: instance labelling of the soft-max of one C2 batch (B*T = 32 frames of
    512x512) on the device.  value: soft-max resident in HBM; e2e: host soft-max in, host uint16 labels out;
    cpu_baseline: the reference's numpy / SciPy / OpenCV algorithm (oracle pinned to the reference's own vectors) on a
    few of the same frames.  HBM-bound integer work: algorithmic bytes = 12 B/px soft-max read + 2 B/px labels."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import __graft_entry__ as ge
    if world > 1:
        if rank == 0:
            ge.build()
        dist.barrier()
    ge.build()
    from lstm_unet_b200.postprocess import PostProcessor
    from lstm_unet_b200 import _lib
    from oracle import postprocess_oracle as P
    n, H, W = args.batch * args.unroll, args.size, args.size
    kw = dict(edge_dist=2, min_cell_size=10, max_cell_size=100, FOV=0)          # CTCInferenceParams defaults
    distinct = [P.synthetic_softmax(H, W, 1000 + 17 * rank + i, 'cells') for i in range(min(n, 8))]
    sm_host = torch.from_numpy(np.stack([distinct[i % len(distinct)] for i in range(n)])).pin_memory()
    sm_dev = sm_host.cuda()
    pp = PostProcessor(**kw)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    labels = pp(sm_dev)
    if rank == 0:       # the checker: first frames against the oracle
        got = labels.numpy()
        for i in range(min(2, n)):
            assert np.array_equal(got[i], P.postprocess_frame(distinct[i], **kw)), 'post-processing differs from the oracle'
    for _ in range(args.warmup):
        pp(sm_dev)
    barrier()
    lib = _lib.load_library()
    lib.lu_post_launch_count(None, 1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        pp(sm_dev)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    import ctypes
    nl = ctypes.c_int64()
    lib.lu_post_launch_count(ctypes.byref(nl), 0)
    clocks = sampler.stop() if rank == 0 else None
    dev_buf = torch.empty_like(sm_dev)

    def e2e_once():
        dev_buf.copy_(sm_host, non_blocking=True)
        return pp(dev_buf).numpy()
    for _ in range(4):          # the pinned read-back ring has three buffers
        e2e_once()
    barrier()
    e2e_steps = max(2, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = e2e_once()
    barrier()
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3
    if world > 1:
        t = torch.tensor([ms, e2e_wall_ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_wall_ms = [float(v) for v in t.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = n * args.steps * world / (ms * 1e-3)
    alg_bytes = n * H * W * 14.0
    hbm = 6550.0
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    src = 'fallback'
    if os.path.exists(pk):
        with open(pk) as f:
            hbm = json.load(f).get('hbm_gbs', hbm)
        src = 'measured hbm_gbs'
    achieved = alg_bytes * args.steps / (ms * 1e-3) / 1e9
    line = {
        'metric': 'frames/sec (512x512 soft-max -> uint16 instance labels)', 'value': value, 'unit': 'frames/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8/i32', 'data': 'synthetic',
        'config': {'workload': 'SURVEY 8f row 3: Inference2D.py:64-123 instance labelling of %d soft-max frames %dx%d per GPU '
                               '(synthetic cell-like maps, ~700 components per frame), CTCInferenceParams defaults' % (n, H, W),
                   'l2_policy': 'soft-max batch + workspace (%.0f MB) exceed the 126 MB L2' % (n * H * W * (12 + 31) / 1e6)},
        'e2e': {'value': n * e2e_steps * world / (e2e_wall_ms * 1e-3), 'unit': 'frames/s',
                'h2d_bytes_per_step': int(sm_host.numel() * 4), 'd2h_bytes_per_step': int(out.nbytes),
                'steps': e2e_steps, 'ms_per_step': e2e_wall_ms / e2e_steps},
        'gpu_launches': int(nl.value), 'clocks': clocks,
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': hbm, 'unit': 'GB/s', 'frac': achieved / hbm,
                     'traffic': None, 'peak_source': src,
                     'kernel': 'whole lu_postprocess pipeline (%d launches per batch); algorithmic bytes = 14 B/pixel'
                               % (nl.value // max(1, args.steps))},
    }
    if world == 1 and not args.no_cpu:
        k = 4
        t0 = time.perf_counter()
        for i in range(k):
            P.postprocess_frame(distinct[i % len(distinct)], **kw)
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': k / dt, 'unit': 'frames/s', 'cores': 1, 'kind': 'port',
                                'sample': 'oracle/postprocess_oracle.postprocess_frame (the reference\'s numpy + SciPy + '
                                          'OpenCV steps, pinned to vectors made by the reference\'s own statements) on %d '
                                          'of the same frames' % k}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_augment(args):
    """--mode augment: the training reader's per-frame augmentation chain (DataHandeling.py:330-380:
    contrast / brightness, cv2.warpAffine + scipy map_coordinates elastic warp of image and segmentation,
    _fix_transformed_segmentation, flips, rot90) for one batch of sequences on the device.  value: crops resident in
    HBM; e2e: host crops in (pinned), device batch out + a 4-byte checksum read back (the batch feeds the model on the
    device); cpu_baseline: the oracle (pinned to vectors made by the reference's helpers) on a few frames."""
    import torch
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import __graft_entry__ as ge
    ge.build()
    from lstm_unet_b200.augment import SequenceAugmenter, random_affine
    from lstm_unet_b200 import _lib
    from oracle import augment_oracle as A
    B, T, H, W = args.batch, args.unroll, args.size, args.size
    rs = np.random.RandomState(7 + rank)
    imgs, segs = A.synthetic_sequence(T, H, W, 11 + rank)
    aug = SequenceAugmenter()
    affine = random_affine((H, W), W * 0.08, rs)
    rand2 = np.stack([rs.rand(H, W), rs.rand(H, W)])
    contrast = (rs.rand(T) + 0.5).astype(np.float32)
    brightness = ((rs.rand(T) - 0.5) * 0.2 * imgs.max()).astype(np.float32)
    coords = aug.elastic_coords(rand2, W * 2, W * 0.15)
    img_pin, seg_pin = torch.from_numpy(imgs).pin_memory(), torch.from_numpy(segs).pin_memory()
    img_dev, seg_dev = img_pin.cuda(), seg_pin.cuda()
    out_i = torch.empty(B * T * H * W, dtype=torch.float32, device='cuda')
    out_s = torch.empty_like(out_i)

    def step(src_i, src_s):
        for b in range(B):          # one launch sequence per sample's sequence chunk, written into its batch slice
            sl = slice(b * T * H * W, (b + 1) * T * H * W)
            aug.augment(src_i, src_s, contrast, brightness, affine, coords, (1, 0), 1, out_img=out_i[sl], out_seg=out_s[sl])
    step(img_dev, seg_dev)
    torch.cuda.synchronize()
    if rank == 0:
        ref_i, ref_s = A.augment_frame(imgs[0], segs[0], contrast[0], brightness[0], affine,
                                       A.elastic_coords(rand2, W * 2, W * 0.15), (1, 0), 1)
        got_s = out_s[:H * W].reshape(H, W).cpu().numpy()
        got_i = out_i[:H * W].reshape(H, W).cpu().numpy()
        assert (got_s != ref_s).mean() < 1e-4 and np.allclose(got_i, ref_i, rtol=1e-5, atol=1e-2), 'augmentation differs from the oracle'
    for _ in range(args.warmup):
        step(img_dev, seg_dev)
    torch.cuda.synchronize()
    lib = _lib.load_library()
    lib.lu_post_launch_count(None, 1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step(img_dev, seg_dev)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    import ctypes
    nl = ctypes.c_int64()
    lib.lu_post_launch_count(ctypes.byref(nl), 0)
    clocks = sampler.stop() if rank == 0 else None
    e2e_steps = max(2, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        img_dev.copy_(img_pin, non_blocking=True)
        seg_dev.copy_(seg_pin, non_blocking=True)
        step(img_dev, seg_dev)
        chk = float(out_i[:16].sum())
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    if rank != 0:
        return
    n = B * T
    alg = n * H * W * 16.0           # image + segmentation read and written once, fp32
    hbm, src = 6550.0, 'fallback'
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        with open(pk) as f:
            hbm = json.load(f).get('hbm_gbs', hbm)
        src = 'measured hbm_gbs'
    ach = alg * args.steps / (ms * 1e-3) / 1e9
    line = {'metric': 'frames/sec (%dx%d crop: elastic + affine + photometric augmentation)' % (H, W), 'value': n * args.steps * world / (ms * 1e-3),
            'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32/f64', 'data': 'synthetic',
            'config': {'workload': 'SURVEY 8f row 4: DataHandeling.py:330-380 augmentation of a batch of %d sequences x %d frames, %dx%d crops, '
                                   'elastic_augmentation=True, randomize=True' % (B, T, H, W),
                       'l2_policy': 'batch + scratch (%.0f MB) exceed the 126 MB L2 at the default size' % (n * H * W * 40 / 1e6)},
            'e2e': {'value': n * e2e_steps * world / (e2e_ms * 1e-3), 'unit': 'frames/s', 'h2d_bytes_per_step': int(imgs.nbytes + segs.nbytes),
                    'd2h_bytes_per_step': 4, 'steps': e2e_steps, 'ms_per_step': e2e_ms / e2e_steps},
            'gpu_launches': int(nl.value), 'clocks': clocks,
            'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s', 'frac': ach / hbm, 'traffic': None,
                         'peak_source': src, 'kernel': 'whole lu_augment_sequence chain (%d launches per batch); algorithmic bytes = 16 B/pixel' % (nl.value // max(1, args.steps))}}
    if world == 1 and not args.no_cpu:
        cref = A.elastic_coords(rand2, W * 2, W * 0.15)
        k = 3
        t0 = time.perf_counter()
        for i in range(k):
            A.augment_frame(imgs[i % T], segs[i % T], contrast[i % T], brightness[i % T], affine, cref, (1, 0), 1)
        dt = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': k / dt, 'unit': 'frames/s', 'cores': 1, 'kind': 'port',
                                'sample': 'oracle/augment_oracle.augment_frame (numpy restatement of the reference chain, pinned to vectors '
                                          'made by the reference\'s helpers; slower than cv2 / scipy themselves) on %d frames' % k}
    emit(line)


def main():
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner) are sent to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='all', choices=['all', 'infer', 'train', 'stream', 'postprocess', 'augment'],
                    help="all (default) = C2 inference headline + a `train` block with the C3/C4 train step; infer / train = one of "
                         "them; stream = Inference2D's real per-frame call (B=1, T=1)")
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'bf16x3', 'fp16'])
    ap.add_argument('--no-parity', dest='no_parity', action='store_true', help='skip the pre-timing spot check against the oracle')
    ap.add_argument('--no-variants', dest='no_variants', action='store_true', help='skip the fp16 / bf16x3 timings of the C2 forward')
    ap.add_argument('--no-full', dest='no_full', action='store_true', help='--impl reference: skip the one step at the true C2 shape and the C1 train step')
    ap.add_argument('--a-mode', dest='a_mode', default='halo', choices=['halo', 'direct'])
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--unroll', type=int, default=8)
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--sync-bn', dest='sync_bn', action='store_true',
                    help='train mode, N>1: BatchNorm statistics over the batch of all ranks (one small fp64 all-reduce per BN layer)')
    ap.add_argument('--allreduce', default='overlapped', choices=['overlapped', 'single'],
                    help='train mode, N>1: gradient exchange started per block from inside the backward, or one collective after it')
    ap.add_argument('--post', action='store_true', help='e2e leg: label every step on the device (postprocess.PostProcessor) and read back the uint16 labels instead of the soft-max (Inference2D.py:59-124)')
    ap.add_argument('--cuda-graph', dest='cuda_graph', default='auto', choices=['auto', 'on', 'off'],
                    help='replay the inference forward as a CUDA graph (auto: launch-bound shapes, B*T <= 2)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.mode == 'stream':
        args.batch, args.unroll = 1, 1
    if args.impl == 'reference':
        run_reference(args)
    elif args.mode == 'postprocess':
        run_postprocess(args)
    elif args.mode == 'augment':
        run_augment(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
