"""Throughput of back-to-back Q1 calls through the public host-buffer API: blocking hw1f_bond_curve, and submit / collect with
2, 3, 4 result slots in flight (every slot is a lane with its own stream: the ramp-up of one call -- jump tables, stream
derivation of the first wave -- overlaps the drain and the tail of another), and two engines on top of that.  Every step uploads its model tables and has P, f, P_se read on the host.
    python tools/two_engine_probe.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hw1f_b200 as hw

N = 1 << 20
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200


def run(engines, depth, n):
    res, pending = None, []
    for i in range(n):
        e = engines[i % len(engines)]
        slot = (i // len(engines)) % (depth // len(engines))
        if len(pending) == depth:
            pe, ps = pending.pop(0)
            res = pe.bond_curve_collect(slot=ps)
        e.set_model(e.params)
        e.bond_curve_submit(hw.Rng(7000 + i, N), slot=slot)
        pending.append((e, slot))
    while pending:
        pe, ps = pending.pop(0)
        res = pe.bond_curve_collect(slot=ps)
    return res


def timed(label, engines, depth):
    run(engines, depth, 6)
    t0 = time.perf_counter()
    res = run(engines, depth, steps)
    ms = (time.perf_counter() - t0) * 1e3 / steps
    print(f"{label:58s} {ms:.4f} ms per step   P(0,10) = {res['P'][-1]:.7f}")
    return res


a, b = hw.Engine(device=0), hw.Engine(device=0)
for e in (a, b):
    e.bond_curve(hw.Rng(1, N))
t0 = time.perf_counter()
for i in range(steps):
    a.set_model(a.params)
    r0 = a.bond_curve(hw.Rng(7000 + i, N), timing=False)
print(f"{'blocking hw1f_bond_curve':58s} {(time.perf_counter() - t0) * 1e3 / steps:.4f} ms per step   P(0,10) = {r0['P'][-1]:.7f}")
rs = [timed("submit / collect, one engine, 2 slots (= lanes) in flight", [a], 2),
      timed("submit / collect, one engine, 3 slots in flight", [a], 3),
      timed("submit / collect, one engine, 4 slots in flight", [a], 4),
      timed("submit / collect, two engines x 4 slots in flight", [a, b], 8)]
assert all((r0["P"] == r["P"]).all() for r in rs)
