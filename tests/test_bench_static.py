"""Static guards on bench.py (it needs a GPU to run, but its collective structure can be checked here)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_no_collective_after_rank_split():
    """every rank must execute the same sequence of all-reduces: nothing that steps the engine
    (device_step / e2e_step contain dist.all_reduce) may come after the non-zero ranks return"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    main = src[src.index("def main():"):]
    split = main.index("if rank != 0:")
    tail = main[split:]
    assert not re.search(r"\b(device_step|e2e_step|all_reduce|barrier)\(", tail), "collective after the rank split"


def test_contract_keys_present():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"',
                '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"', '"e2e"', '"h2d_bytes_per_step"',
                '"d2h_bytes_per_step"', '"gpu_launches"', '"clocks"', '"roofline"', '"cpu_baseline"', '"impl"'):
        assert key in src, key
