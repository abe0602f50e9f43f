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
    # the engine launches on ITS stream (its own non-blocking one unless the caller passed stream=...), the
    # collective runs on torch's current stream: order them explicitly on both sides
    torch.cuda.current_stream().synchronize()      # the zero-fill above has landed
    engine.bond_curve_moments(rng_factory(first, n), moments.data_ptr())
    engine.synchronize()                           # the moments are written before NCCL reads them
    allreduce_moments(moments)
    torch.cuda.current_stream().synchronize()      # ... and reduced before the engine's stream finalises them
    return engine.bond_curve_finish(moments.data_ptr(), n_total)


class PeerAllReduce:
    """hw1f_comm_*: the moment all-reduce as one own kernel over NVLink peer memory (CUDA IPC mailboxes).
    torch.distributed is only used once, to exchange the 64-byte IPC handles."""

    def __init__(self, engine, stream):
        import ctypes as C
        from . import _ffi
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerAllReduce needs an initialised process group")
        self._C, self._lib = C, _ffi.load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        handle = (C.c_ubyte * 64)()
        h = C.c_void_p()
        st = self._lib.hw1f_comm_create(engine._h, self.world, handle, C.byref(h))
        if st != _ffi.OK:
            raise RuntimeError(f"hw1f_comm_create failed with status {st}")
        self._h = h
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device="cuda")
        allh = torch.empty(64 * self.world, dtype=torch.uint8, device="cuda")
        dist.all_gather_into_tensor(allh, mine)
        blob = bytes(allh.cpu().tolist())
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        st = self._lib.hw1f_comm_connect(self._h, self.rank, buf, C.c_void_p(stream.cuda_stream))
        if st != _ffi.OK:
            msg = self._lib.hw1f_comm_last_error(self._h).decode()
            self._lib.hw1f_comm_destroy(self._h)
            self._h = None
            raise RuntimeError("hw1f_comm_connect: " + msg)
        self.attached = False
        dist.barrier()

    def all_reduce(self, moments):
        """in-place rank-ordered SUM of a float64 CUDA tensor with <= 512 elements (async, stream ordered)"""
        if moments.dtype != torch.float64 or moments.numel() > 512:
            raise TypeError("float64 tensor with at most 512 elements")
        st = self._lib.hw1f_comm_allreduce(self._h, self._C.c_void_p(moments.data_ptr()), moments.numel())
        if st != 0:
            raise RuntimeError(self._lib.hw1f_comm_last_error(self._h).decode())
        return moments

    def attach(self, on=True):
        """on: every engine.*_moments call from now on returns the ALL-REDUCED vector -- the last block of the
        simulation kernel's reduction posts it to the peers itself (no separate collective launch).  Every rank
        must make the same calls in the same order."""
        st = self._lib.hw1f_comm_attach(self._h, 1 if on else 0)
        if st != 0:
            raise RuntimeError(f"hw1f_comm_attach failed with status {st}")
        self.attached = bool(on)
        return self

    def timeouts(self):
        n = self._C.c_uint32()
        self._lib.hw1f_comm_timeouts(self._h, self._C.byref(n))
        return n.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hw1f_comm_destroy(self._h)
            self._h = None
