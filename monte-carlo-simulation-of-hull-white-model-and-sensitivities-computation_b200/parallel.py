"""Multi-GPU plumbing: one process per GPU (torchrun), paths sharded by XORWOW subsequence range,
one all-reduce of the packed double moment vector per workload.

Path p's randomness is XORWOW subsequence p, so the union of the rank shards is bit-identical to
the single-GPU path set; nothing but the moment vector (a few hundred doubles) ever crosses NVLink.
"""
import torch
import torch.distributed as dist


def shard_paths(n_total, rank, world):
    """contiguous split of [0, n_total) into `world` ranges; returns (first_path, n_paths).
    The first n_total % world ranks get one extra path; ranges need no alignment (the kernels
    mask ragged chunk edges)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_total), int(world))
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def allreduce_moments(moments):
    """in-place SUM all-reduce of a float64 moment tensor (NCCL on GPUs, gloo in the CPU tests);
    a no-op without an initialised process group"""
    if moments.dtype != torch.float64:
        raise TypeError("moment vectors are float64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(moments, op=dist.ReduceOp.SUM)
    return moments


def sharded_bond_curve(engine, rng_factory, n_total, device=None):
    """Q1 across the process group: every rank simulates its shard, moments are all-reduced, every
    rank finalises redundantly.  rng_factory(first_path, n_paths) -> Rng."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    first, n = shard_paths(n_total, rank, world)
    moments = torch.zeros(2 * engine.n_mat, dtype=torch.float64, device=device or "cuda")
    torch.cuda.current_stream().synchronize()
    engine.bond_curve_moments(rng_factory(first, n), moments.data_ptr())
    allreduce_moments(moments)
    return engine.bond_curve_finish(moments.data_ptr(), n_total)
