"""Multi-GPU plumbing: one process per GPU (torch.distributed), the frame sharded by SAMPLE RANGE
 and the per-GPU float accumulation buffers combined with ONE sum-reduce to rank 0.
The Philox key is (pixel, sample, bounce), so the set of sample contributions is identical for
any number of ranks; only the float summation order of the reduce differs."""


def sample_range(rank, world, spp, spp_begin=0):
    """Rank r of `world` renders samples [begin + r*spp//world, begin + (r+1)*spp//world)."""
    if not (0 <= rank < world) or spp < 0:
        raise ValueError("bad rank/world/spp")
    return spp_begin + rank * spp // world, spp_begin + (rank + 1) * spp // world


def reduce_accumulators(accum, dist=None, dst=0):
    """Sum-reduce the accumulation buffer (a torch tensor on this rank's device) to rank `dst`.
    NCCL over NVLink on GPUs, gloo in the CPU tests.  No-op for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return accum
    dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum
