"""Multi-GPU plumbing for the path: one process per GPU, sequences sharded by batch.

Every sample's frames and its stateful h/c at all levels live on exactly one rank, so the forward needs no data-path
collective.  The training step has one exchange: a sum all-reduce of the flat gradient buffer (NCCL over NVLink on the
GPU box; the same code runs over gloo in the CPU tests).  Timing is the max over ranks."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world):
    """Samples [lo, hi) owned by `rank`.  The global batch must divide evenly (each rank's state shape is frozen)."""
    if global_batch % world:
        raise ValueError('global batch %d is not divisible by %d ranks' % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def all_reduce_mean_(flat):
    """In-place mean all-reduce of a flat gradient tensor (one collective per step)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == 'nccl':
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)          # one pass: NCCL averages inside the collective
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)          # gloo (CPU tests) has no AVG
            flat.div_(dist.get_world_size())
    return flat


def max_over_ranks(value, device='cpu'):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(value)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(value)


class OverlappedAllReduce:
    """Mean all-reduce of the flat gradient buffer, started block by block from inside the backward.

    ``lu_loss_backward`` reports every Up / Down block whose gradient range is final (decoder first); each range is
    all-reduced on a side stream ordered after the compute stream at that point, so the exchange of the large bottleneck
    block overlaps the backward of the encoder.  Pass an instance as ``allreduce=`` to ``ULSTMnet2D.train_step``: it is
    armed with ``begin(session, grads)`` before the backward and called like ``all_reduce_mean_`` after it (which then only
    joins the side stream).  On CPU tensors (gloo, the tests) the ranges are reduced synchronously."""

    def __init__(self):
        self.side = None
        self.ranges = []
        self._grads = None

    def active(self):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def begin(self, session, grads):
        self._grads = grads
        self.ranges = []
        if not self.active():
            session.set_grad_bucket_callback(None)
            return
        if grads.is_cuda and self.side is None:
            self.side = torch.cuda.Stream(device=grads.device)
        session.set_grad_bucket_callback(self._on_bucket)

    def _on_bucket(self, offset, count):
        g = self._grads[offset:offset + count]
        self.ranges.append((offset, count))
        if g.is_cuda:
            self.side.wait_stream(torch.cuda.current_stream(g.device))
            with torch.cuda.stream(self.side):
                all_reduce_mean_(g)
        else:
            all_reduce_mean_(g)

    def __call__(self, flat):
        if not self.active():
            return flat
        covered = sum(c for _, c in self.ranges)
        if not self.ranges:                    # no bucket was reported (callback not armed): one collective, as before
            all_reduce_mean_(flat)
        elif covered != flat.numel():
            raise RuntimeError('gradient buckets cover %d of %d elements' % (covered, flat.numel()))
        if flat.is_cuda and self.side is not None:
            torch.cuda.current_stream(flat.device).wait_stream(self.side)
        return flat


def enable_sync_batchnorm(session):
    """Synchronised BatchNorm (SURVEY 8e option ii): every BN layer of `session` normalises with the statistics of the
    batch of ALL ranks, like the reference does on one device.  Adds one small float64 sum all-reduce per BN layer to the
    training forward and one to the backward (enqueued on the compute stream by the callback).  Every rank must hold the
    same number of frames.  No-op without an initialised process group of more than one rank."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        session.set_bn_sync_callback(None)
        return False

    def sum_over_ranks(ptr, count):
        v = session.workspace_view(ptr, count, np.float64)
        t = torch.from_numpy(v) if not isinstance(v, torch.Tensor) else v
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    session.set_bn_sync_callback(sum_over_ranks, dist.get_world_size())
    return True
