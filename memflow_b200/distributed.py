"""Multi-GPU sharding of the importance-sampling integral.

Events are independent, so the global event range [0, N_total) is split contiguously, one process per
GPU; the Philox counter is the GLOBAL event index, so the estimator's inputs do not depend on the number
of ranks. The only collective is one all-reduce of the 3 accumulators (sum w, sum w^2, n_valid) at the end
-- never per chunk. The reference itself has no explicit collective (DDP only).
"""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world_size):
    """Contiguous [begin, end) owned by `rank`; the union over ranks is exactly [0, n_total)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return (n_total * rank) // world_size, (n_total * (rank + 1)) // world_size


def chunk_ranges(begin, end, chunk):
    """Split [begin, end) into consecutive chunks of at most `chunk` events."""
    out = []
    b = begin
    while b < end:
        e = min(b + chunk, end)
        out.append((b, e))
        b = e
    return out


def allreduce_accumulators(acc, group=None):
    """In-place SUM all-reduce of the double[3] accumulator (NCCL on GPUs, gloo in the CPU tests).
    n_valid travels as a double: exact up to 2^53 events."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


class ShardedIntegrator:
    """Runs an `ImportanceSampler` over this rank's shard of [0, n_total) in chunks and reduces once.

    `phi_provider(begin, end)` returns the float32 [T, end-begin, F, 3K-1] spline parameters for the
    global events [begin, end) (in MEMFlow: the flow's hyper-network output; in the benchmark: a resident
    synthetic buffer)."""

    def __init__(self, sampler, n_total, chunk, rank=0, world_size=1, group=None):
        self.sampler = sampler
        self.n_total = int(n_total)
        self.chunk = int(chunk)
        self.rank, self.world_size, self.group = rank, world_size, group
        self.begin, self.end = shard_range(self.n_total, rank, world_size)

    def run(self, phi_provider, acc=None):
        from .estimator import new_accumulator
        if acc is None:
            acc = new_accumulator(self.sampler.device)
        for b, e in chunk_ranges(self.begin, self.end, self.chunk):
            self.sampler.run_chunk(phi_provider(b, e), event_offset=b, acc=acc)
        return allreduce_accumulators(acc, self.group)
