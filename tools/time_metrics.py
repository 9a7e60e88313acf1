"""Device time of the per-step metrics (lu_seg_measure: SEG measure + accuracy, train2D.py:97-102) at the C3 step shape,
next to the CPU oracle (the reference's algorithm) on a few frames.  Prints one JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

ge.build()
from lstm_unet_b200 import losses  # noqa: E402
from oracle import seg_oracle as S  # noqa: E402

B, T, H, W = 4, 8, 512, 512
labels, logits = S.synthetic_pair(B, T, H, W, 5, 'blobs')
lab, lg = torch.from_numpy(labels).cuda(), torch.from_numpy(logits).cuda()
calc = losses.seg_measure(2)
got = calc(lab, lg)
t0 = time.perf_counter()
want = S.seg_measure(labels[:1, :2], logits[:1, :2])
cpu_s_per_frame = (time.perf_counter() - t0) / 2
assert abs(calc(lab[:1, :2].contiguous(), lg[:1, :2].contiguous()) - want) < 1e-6 * max(1.0, abs(want))
for _ in range(3):
    calc(lab, lg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 50
for _ in range(n):
    calc(lab, lg)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({'what': 'SEG measure + accuracy of one C3 step (32 frames 512x512), device incl. the 32-byte read-back',
                  'ms_per_step': ms, 'frames_per_s': B * T / (ms * 1e-3), 'seg': got, 'accuracy': calc.last_accuracy,
                  'cpu_oracle_frames_per_s': 1.0 / cpu_s_per_frame, 'cpu_cores': 1}))
