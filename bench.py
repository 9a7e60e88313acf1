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


def lstm_traffic_bytes():
    """DRAM bytes of one level-1 ConvLSTM launch from the committed `ncu --set full` summary (profiles/), or None."""
    p = os.path.join(ROOT, 'profiles', 'r1_ncu_prof_lstm_l1.txt')
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
        q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
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


def cpu_reference_run(steps, warmup, B=1, T=4, H=256, W=256):
    """CPU restatement of the reference forward (Inference2D call: pad_image=True, training=False) on a bounded
    sample of the workload (smaller frames / batch, same network, same T-unrolled stateful call)."""
    import torch
    from oracle import lstm_unet_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    net = O.OracleNet(O.CTC_NET_PARAMS, 'NCHW', True, seed=0)
    x = torch.randn(B, T, 1, H, W)
    with torch.no_grad():
        for _ in range(warmup):
            net(x, False)
        t0 = time.perf_counter()
        for _ in range(steps):
            net(x, False)
        dt = time.perf_counter() - t0
    fps = B * T * steps / dt
    # frames differ in size from the workload's: scale by pixels so the number is frames/s of 512x512 frames
    scale = (H * W) / (512.0 * 512.0)
    return {'value': fps * scale, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
            'sample': 'oracle (torch-CPU restatement; TF not installable) forward, CTC net, B=%d T=%d %dx%d pad_image, '
                      '%d steps; %.3f frames/s at %dx%d scaled by pixel count to 512x512' % (B, T, H, W, steps, fps, H, W),
            'ms_per_step': dt / steps * 1e3}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    cb = cpu_reference_run(steps, max(1, min(args.warmup, 1)))
    line = {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': steps, 'warmup': 1, 'ms_per_step': cb['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'C2 inference forward 512x512 T=8 B=4 (bounded sample, see cpu_baseline.sample)'},
            'cpu_baseline': {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': cb['value'], 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    emit(line)


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

    B, T, H, W = args.batch, args.unroll, args.size, args.size
    training = args.mode == 'train'
    model = ULSTMnet2D(CTC_NET, 'NCHW', pad_image=not training, precision=args.precision, a_mode=args.a_mode,
                       seed=0 if training else rank, train=training,
                       cuda_graph={'auto': 'auto', 'on': True, 'off': False}[args.cuda_graph],
                       sync_bn=bool(training and args.sync_bn))
    rng = np.random.default_rng(1234 + rank)
    x_host = rng.standard_normal((B, T, 1, H, W)).astype(np.float32)
    x_dev = torch.from_numpy(x_host).cuda()
    if training:
        # config 3/4: full train step = forward(training=True) + weighted CE + backward + (N>1: one NCCL all-reduce of
        # the flat gradients) + Keras Adam; same initial weights on every rank (seed 0), batch-sharded data
        from lstm_unet_b200.Networks import Adam
        from lstm_unet_b200.parallel import all_reduce_mean_, OverlappedAllReduce
        reducer = None
        if world > 1:
            reducer = OverlappedAllReduce() if args.allreduce == 'overlapped' else all_reduce_mean_
        lab_host = rng.integers(-1, 3, size=(B, T, 1, H, W)).astype(np.float32)
        lab_dev = torch.from_numpy(lab_host).cuda()
        opt = Adam(lr=1e-5)
        cw = [0.15, 0.25, 0.6]
        fwd = model

        class _Step:
            def __call__(self, x, tr):
                lab = lab_dev if x.__class__ is not np.ndarray else lab_host
                sm, lg, loss = fwd.train_step(x, lab, cw, opt, reducer)
                return loss, sm
        step_fn = _Step()
    else:
        step_fn = None
    run = (lambda x, tr: step_fn(x, tr)) if training else (lambda x, tr: model(x, tr))
    # shard by batch: every rank owns its B samples and their recurrent states (weak scaling, no data-path collective)
    run(x_dev, training)
    torch.cuda.synchronize()
    sess = model._sess
    flops_step = sess.forward_flops(T) * B * (3 if training else 1)     # train step = 3 x forward
    lstm_flops_step = sess.lstm_flops(T) * B

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        run(x_dev, training)
    barrier()
    sess.launch_count(reset=True)
    sess.lstm_kernel_time(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        run(x_dev, training)          # states carry over between iterations, like the train / inference loops
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = sess.launch_count()
    lstm_ms, lstm_n = sess.lstm_kernel_time(False)
    clocks = sampler.stop() if rank == 0 else None

    # end to end through the public API with host buffers: pinned H2D of the frames + D2H of the soft-max
    post = None
    if args.post and not training:
        from lstm_unet_b200.postprocess import PostProcessor
        post = PostProcessor()

    def e2e_once():
        out = run(x_host, training)
        if post is not None:
            return post(out[1]).numpy()
        # inference: D2H of the soft-max (Inference2D.py:60); training: D2H of the loss (train2D.py:103 -> metrics)
        return out[1].numpy() if not training else np.asarray(float(out[0]), dtype=np.float32)
    for _ in range(4):          # >= 3: the .numpy() read-back cycles through a ring of three pinned host buffers, each
        e2e_once()              # allocated (cudaHostAlloc, tens of ms for 100 MB) the first time it is used
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        sm = e2e_once()
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3 * 0.0)
    e2e_wall_ms = (time.perf_counter() - t0) * 1e3

    if world > 1:
        t = torch.tensor([ms, e2e_ms, e2e_wall_ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, e2e_wall_ms = [float(v) for v in t.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    frames = B * T * args.steps * world
    value = frames / (ms * 1e-3)
    e2e_value = B * T * e2e_steps * world / (e2e_wall_ms * 1e-3)
    sustained, burst, which = peaks()
    lstm_tflops = (lstm_flops_step * args.steps / (lstm_ms * 1e-3)) / 1e12 if lstm_ms > 0 else None
    cb = cpu_reference_run(4, 1) if (world == 1 and not args.no_cpu) else None      # ~10 s of host work on 16 cores
    line = {
        'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'bf16x3(split-bf16, fp32-equivalent)', 'data': 'synthetic',
        'config': {'workload': '%s: ConvLSTM-UNet (CTCParams net, 74.6M params) %s, %dx%d, T=%d, batch %d per GPU, '
                               'pad_image=%s, stateful' % ('C3' if training else 'C2', 'full train step (fwd+loss+bwd+Adam)' if training
                                                           else 'inference forward', H, W, T, B, not training),
                   'global_batch': B * world, 'parallelism': ('data parallel x%d: batch-sharded, NCCL mean all-reduce of the 74.6M fp32 gradients per step (started per block from inside the backward)'
                                   if training else 'batch-sharded replicas x%d (no data-path collective)') % world,
                   'l2_policy': 'inputs+activations per step (>6 GB) exceed the 126 MB L2; no flush needed',
                   'a_mode': args.a_mode, 'step_tflop': flops_step / 1e12, 'cuda_graph': bool(model.graph_active),
                   # experiment switches of the library that were set for this run (none = the default kernels)
                   'switches': {k: os.environ[k] for k in ('LU_PAIR', 'LU_WGRAD_CLUSTER', 'LU_CLUSTER', 'LU_CLUSTER_WIDE',
                                                           'LU_B_RESIDENT', 'LU_WGRAD_ENGINE') if k in os.environ}},
        'e2e': {'value': e2e_value, 'unit': 'frames/s',
                'h2d_bytes_per_step': int(x_host.nbytes) * (2 if training else 1),
                'd2h_bytes_per_step': int(sm.nbytes), 'steps': e2e_steps, 'ms_per_step': e2e_wall_ms / e2e_steps,
                'labelled_on_device': post is not None},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {'bound': 'tensor', 'achieved': lstm_tflops, 'peak': sustained, 'unit': 'TFLOP/s',
                     'frac': (lstm_tflops / sustained) if lstm_tflops else None, 'traffic': lstm_traffic_bytes(),
                     'traffic_note': 'DRAM bytes of one level-1 ConvLSTM launch (ncu --set full, profiles/r1_ncu_prof_lstm_l1.txt)',
                     'kernel': 'lu_conv_tc_kernel<LSTM> forward launches (all 4 ConvLSTM levels, %d launches)' % lstm_n,
                     'kernel_ms_per_step': lstm_ms / args.steps, 'kernel_share_of_step': lstm_ms / ms if ms else None,
                     'peak_source': which + ' bf16_tflops_sustained (kernel timed inside a long step); burst %.1f' % burst,
                     'whole_step_tflops': flops_step * args.steps / (ms * 1e-3) / 1e12},
    }
    if cb is not None:
        line['cpu_baseline'] = {k: cb[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
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
    ap.add_argument('--mode', default='infer', choices=['infer', 'train', 'stream', 'postprocess', 'augment'],
                    help="infer = C2 (default, headline); train = C3/C4 full train step; stream = Inference2D's real per-frame call (B=1, T=1)")
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'bf16x3'])
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
