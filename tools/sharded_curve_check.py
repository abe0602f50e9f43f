"""torchrun --nproc-per-node N tools/sharded_curve_check.py
parallel.sharded_bond_curve with a DEFAULT-constructed Engine (own non-blocking stream, not torch's current
stream): the helper must order the engine's stream against the collective's stream itself.  Every rank compares
the sharded result with the whole path set simulated alone.  Exit code 0 = all ranks agree."""
import datetime
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hw1f_b200 as hw  # noqa: E402


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    eng = hw.Engine(device=local)                      # default: the engine's own stream
    n_total, seed = (1 << 18) + 3, 99
    worst = 0.0
    for it in range(20):                               # repeated: a race shows up as an occasional mismatch
        got = hw.package.parallel.sharded_bond_curve(
            eng, lambda first, n: hw.Rng(seed + it, n, first_path=first), n_total, device=f"cuda:{local}")
        one = eng.bond_curve(hw.Rng(seed + it, n_total))
        worst = max(worst, float(np.abs(got["P"] / one["P"] - 1.0).max()), float(np.abs(got["f"] - one["f"]).max()))
    ok = torch.tensor([1 if worst < 2e-6 else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"sharded_bond_curve {'ok' if int(ok.item()) else 'MISMATCH'} on {world} GPUs: worst {worst:.3e}")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(ok.item()) else 1)


if __name__ == "__main__":
    main()
