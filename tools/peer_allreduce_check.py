"""torchrun --nproc-per-node N tools/peer_allreduce_check.py
Checks hw1f_comm_* (own NVLink peer-memory all-reduce kernel) against torch.distributed/NCCL on random
vectors and on real moment vectors, and times both.  Exit code 0 = all ranks agree."""
import datetime
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hw1f_b200 as hw  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng = hw.Engine(device=local, stream=stream.cuda_stream)
    peer = hw.package.parallel.PeerAllReduce(eng, stream)
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    worst = 0.0
    for it in range(200):
        n = [202, 5, 256, 1, 10, 212, 404, 512][it % 8]
        x = torch.randn(n, dtype=torch.float64, device="cuda", generator=g) * (10.0 ** (it % 7))
        a, b = x.clone(), x.clone()
        peer.all_reduce(a)
        dist.all_reduce(b)
        torch.cuda.synchronize()
        worst = max(worst, float(((a - b).abs() / (b.abs() + 1e-300)).max()))
        # bit-identical on every rank (rank-ordered sum)
        ref = a.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, a), "peer all-reduce differs between ranks"
    # real moment vectors
    n_paths = 1 << 16
    m1 = torch.zeros(202, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    eng.bond_curve_moments(hw.Rng(99, n_paths, first_path=rank * n_paths), m1.data_ptr())
    m2 = m1.clone()
    peer.all_reduce(m1)
    dist.all_reduce(m2)
    torch.cuda.synchronize()
    assert torch.allclose(m1, m2, rtol=1e-13, atol=0), (m1 - m2).abs().max()
    # timing: K back-to-back all-reduces of 202 doubles
    x = torch.ones(202, dtype=torch.float64, device="cuda")
    res = {}
    for name, fn in (("peer", lambda: peer.all_reduce(x)), ("nccl", lambda: dist.all_reduce(x))):
        for _ in range(20):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn()
            x.fill_(1.0)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 200 * 1e3
    assert worst < 1e-13, worst
    # attached: the tail kernel behind every *_moments simulation exchanges the vector itself (hw1f_comm_attach);
    # curve (202 doubles), ZBC (5), pathwise (2) and the fused vector (220) against NCCL on the same shard moments
    mkt = eng.bond_curve(hw.Rng(1234, 1 << 16))
    cases = (
        (202, lambda r, m: eng.bond_curve_moments(r, m.data_ptr())),
        (5, lambda r, m: eng.zbc_cv_moments(r, mkt["P"], mkt["f"], m.data_ptr(), n_steps_S1=500)),
        (2, lambda r, m: eng.vega_pathwise_moments(r, mkt["P"], mkt["f"], m.data_ptr(), n_steps_S1=500)),
        (220, lambda r, m: eng.fused_moments(r, mkt["P"], mkt["f"], m.data_ptr(), eps=0.001, n_steps_S1=500)),
    )
    for n, sim in cases:
        local = torch.zeros(n, dtype=torch.float64, device="cuda")
        sim(hw.Rng(99, n_paths, first_path=rank * n_paths), local)
        torch.cuda.synchronize()
        dist.all_reduce(local)
        peer.attach(True)
        tail = torch.zeros(n, dtype=torch.float64, device="cuda")
        sim(hw.Rng(99, n_paths, first_path=rank * n_paths), tail)
        peer.attach(False)
        torch.cuda.synchronize()
        assert torch.allclose(tail, local, rtol=1e-13, atol=0), (n, (tail - local).abs().max())
        ref = tail.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, tail), "tail exchange differs between ranks"
    assert peer.timeouts() == 0
    if rank == 0:
        print(f"peer all-reduce ok on {world} GPUs (stand-alone kernel and tail exchange): max rel diff vs NCCL {worst:.2e}; "
              f"us per 202-double all-reduce: peer {res['peer']:.1f}, nccl {res['nccl']:.1f}", flush=True)
    peer.close()
    eng.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
