"""Multi-GPU plumbing for the path: one process per GPU, sequences sharded by batch.

Every sample's frames and its stateful h/c at all levels live on exactly one rank, so the forward needs no data-path
collective.  The training step has one exchange: a sum all-reduce of the flat gradient buffer (NCCL over NVLink on the
GPU box; the same code runs over gloo in the CPU tests).  Timing is the max over ranks."""
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
