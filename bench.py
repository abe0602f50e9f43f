#!/usr/bin/env python
"""bench.py -- HW1F path-steps/s on the BASELINE.json headline workload.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: the Q1 bond-curve workload of
BASELINE.json (`configs[0]`, the configuration the metric is quoted on): 2^20 RNG subsequences x 2
antithetic paths x 1000 exact-discretisation steps per GPU, stateless XORWOW seeding included,
deterministic two-level reduction to 202 double moments.  With N > 1 every rank simulates its own
disjoint subsequence range (weak scaling) and the moment vectors are combined with ONE NCCL
all-reduce of 202 doubles per step.

`value`  : device-timed (CUDA events on the launching stream), nothing crosses PCIe in the region.
`e2e`    : the public host-buffer call a user makes (set_model H2D + simulate + finalise + D2H of
           P, f, P_se), wall-clock between synchronisations.
`roofline`: this path is instruction-issue / FP32+XU pipe bound, not HBM or tensor:
           achieved = algorithmic pipe instructions/s (12 issue slots per path-step), peak = the
           issue rate measured by the engine's own pipe probes on this GPU at the clock seen.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "HW1F path-steps/s (2M antithetic paths, 1000 steps)"
UNIT = "path-steps/s"
N_PATHS_LOG2 = 20          # reference N_PATHS = 1024*1024 (include/common.cuh:16)
# Algorithmic pipe instructions per path-step.
#   reference-order arithmetic: 6.75 FP32 + 1.0 MUFU + 3.75 INT + 0.5 I2F = 12 issue slots
#   decomposed arithmetic (default mode)   : 2.25 FP32 + 1.0 MUFU + 3.75 INT + 0.5 I2F = 7.5 issue slots
#     (per Box-Muller pair and lane = 4 path-steps: 4 Box-Muller FP32 + 5 for the two-step recursion; the loop
#      body's SASS is 299 dispatch cycles per 40 path-steps = 7.48)
#   Q1 save points: reference-order mode adds 4 MUFU.EX2 per 40 path-steps (1.1 MUFU per path-step); decomposed
#   mode evaluates 2cosh(z)-2 as a polynomial on the FMA pipe (no XU work)
ALGO = {
    "decomposed": {"issue": 7.5, "fp32": 2.25, "xu": 1.0},
    "reference_order": {"issue": 12.0, "fp32": 6.75, "xu": 1.1},
}


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons with NVML while the timed regions run"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop = index, [], threading.Event()
        self.max_mhz, self.ok = None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(
                    nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.samples.append((time.time(), mhz, reasons, power))
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._stop.set()

    def summary(self, t0, t1):
        sel = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        names = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
                 0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
                 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}
        bits = 0
        for s in sel:
            bits |= s[2]
        reasons = [n for b, n in names.items() if bits & b and n != "gpu_idle"]
        return {"sm_mhz": statistics.median(s[1] for s in sel), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "power_w_max": max(s[3] for s in sel), "samples": len(sel)}


def run_reference_arm(args, rank, real_stdout):
    """the reference's own CUDA implementation of the path (the reference has no CPU path), driven
    by oracle/_ref/ref_harness on GPU 0; falls back to the OpenMP oracle port if it was not built"""
    if rank != 0:
        return
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    n_paths = 1 << N_PATHS_LOG2
    line = {"metric": METRIC, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Q1 bond curve: 2^20 subsequences x 2 antithetic x 1000 steps (reference "
                                   "simulate_zcb), seeds pinned",
                       "parallelism": "single GPU on rank 0 (the reference has no multi-GPU path)"}}
    if os.path.exists(harness):
        out = os.path.join(ROOT, "gpurun_out", "ref_bench_q1.json")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
        subprocess.run([harness, "bench", "q1", str(args.steps), str(args.warmup), out], check=True, env=env,
                       stdout=subprocess.DEVNULL, timeout=1800)
        with open(out) as f:
            r = json.load(f)
        ms = r["workload_ms_per_step"]
        value = r["path_steps_per_step"] / (ms * 1e-3)
        line.update({"value": value, "ms_per_step": ms, "gpu_launches": 3 * args.steps,
                     "kernel_only": {"value": r["path_steps_per_step"] / (r["kernel_ms_per_step"] * 1e-3),
                                     "ms_per_step": r["kernel_ms_per_step"],
                                     "note": "simulate_zcb alone: the reference's own published metric "
                                             "(src/1_bond_pricing.cu:64-71), init_rng excluded"},
                     "init_rng_ms_per_step": r["init_rng_ms_per_step"],
                     "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                                      "sample": "unmodified reference kernels rebuilt for sm_100 (CUDA-only reference, "
                                                "no CPU path): init_rng + simulate_zcb + compute_average_and_forward "
                                                "+ D2H per step on the same B200, CUDA-event timed"},
                     "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    else:
        from oracle_lib import Oracle
        o = Oracle()
        n = 1 << 14
        t0 = time.time()
        for _ in range(max(args.steps, 1)):
            o.bond_curve(1234, n)
        dt = (time.time() - t0) / max(args.steps, 1)
        value = 2.0 * n * 1000 / dt
        line.update({"value": value, "ms_per_step": dt * 1e3, "gpu_launches": 0,
                     "cpu_baseline": {"value": value, "unit": UNIT, "cores": o.max_threads(), "kind": "port",
                                      "sample": "OpenMP oracle, 2^14 pairs x 1000 steps per step"},
                     "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    _emit(line, real_stdout)


def _emit(line, real_stdout):
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def main():
    # stdout carries exactly ONE JSON line (rank 0): everything libraries print (e.g. NCCL's
    # version banner) is diverted to stderr while the benchmark runs
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--paths-log2", type=int, default=N_PATHS_LOG2, help="subsequences per GPU (default 2^20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="moment all-reduce for N > 1: own NVLink peer-memory kernel (hw1f_comm_*) or NCCL")
    ap.add_argument("--mode", default="decomposed", choices=["decomposed", "reference_order"],
                    help="simulation arithmetic (include/hw1f.h HW1F_MODE_*); both give the same Gaussians")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, real_stdout)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import hw1f_b200 as hw

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the HW1F engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a short watchdog: a mismatched collective should fail in minutes, not hang the box
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # a dedicated non-default stream: the engine launches on it, torch events / NCCL see it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng = hw.Engine(device=local_rank, stream=stream.cuda_stream)
    eng.set_mode(hw._ffi.MODE_DECOMPOSED if args.mode == "decomposed" else hw._ffi.MODE_REFERENCE_ORDER)
    n_paths = 1 << args.paths_log2
    n_steps, n_mat = eng.n_steps, eng.n_mat
    first_path = rank * n_paths                       # disjoint XORWOW subsequence ranges per rank
    path_steps_per_step = 2.0 * n_paths * n_steps * world
    moments = torch.zeros(2 * n_mat, dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    # the single collective of the path: 202 doubles.  Default: the engine's own peer-memory kernel over
    # NVLink (rank-ordered, bit-identical on all ranks); NCCL if asked for or if CUDA IPC is unavailable.
    peer, collective = None, "none"
    if world > 1:
        collective = "nccl"
        want_peer = torch.tensor([1 if args.collective == "peer" else 0], device=dev)
        if args.collective == "peer":
            try:
                peer = hw.package.parallel.PeerAllReduce(eng, stream)
            except Exception as exc:                   # noqa: BLE001
                print(f"[rank {rank}] peer all-reduce unavailable ({exc}); using NCCL", file=sys.stderr)
                want_peer[0] = 0
        dist.all_reduce(want_peer, op=dist.ReduceOp.MIN)   # all ranks or none
        if int(want_peer.item()) == 1:
            collective = "peer_nvlink_kernel"
        elif peer is not None:
            peer.close()
            peer = None

    def reduce_moments():
        if peer is not None:
            peer.all_reduce(moments)
        elif world > 1:
            dist.all_reduce(moments)

    def device_step(seed):
        rng = hw.Rng(seed, n_paths, first_path=first_path)
        eng.bond_curve_moments(rng, moments.data_ptr())
        reduce_moments()

    def e2e_step(seed):
        eng.set_model(eng.params)                    # compute_constants(): H2D of the model tables
        rng = hw.Rng(seed, n_paths, first_path=first_path)
        eng.bond_curve_moments(rng, moments.data_ptr())
        reduce_moments()
        return eng.bond_curve_finish(moments.data_ptr(), n_paths * world)   # D2H of P, f, P_se

    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---- warm-up (also builds the jump tables once; they are seed independent) ----
    for i in range(args.warmup):
        device_step(1000 + i)
        flush.zero_()
    barrier()

    # ---- device-timed region: per-step event pairs, L2 flushed between steps ----
    launches0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_load0 = time.time()
    for i in range(args.steps):
        ev[i][0].record()
        device_step(5000 + i)
        ev[i][1].record()
        flush.zero_()
    barrier()
    t_load1 = time.time()
    launches = eng.launch_count - launches0 + (args.steps if peer is not None else 0)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(total_ms.item()) / args.steps
    value = path_steps_per_step / (ms_per_step * 1e-3)

    # ---- end-to-end region: public host-buffer API, wall clock between synchronisations ----
    for i in range(3):
        e2e_step(9000 + i)
    barrier()
    t0 = time.time()
    for i in range(args.steps):
        last = e2e_step(7000 + i)
    barrier()
    e2e_s = torch.tensor([time.time() - t0], dtype=torch.float64, device=dev)
    t_load2 = time.time()
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_s.item()) * 1e3 / args.steps
    e2e_value = path_steps_per_step / (e2e_ms * 1e-3)
    sampler.stop()
    clocks = sampler.summary(t_load0, t_load2)

    collective_check = None
    if peer is not None:   # same local moments through both collectives (every rank takes part)
        eng.bond_curve_moments(hw.Rng(424242, n_paths, first_path=first_path), moments.data_ptr())
        via_nccl = moments.clone()
        peer.all_reduce(moments)
        dist.all_reduce(via_nccl)
        torch.cuda.synchronize()
        rel = float(((moments - via_nccl).abs() / (via_nccl.abs() + 1e-300)).max())
        collective_check = {"max_rel_diff_vs_nccl": rel, "timeouts": peer.timeouts()}

    # the same steps in the other arithmetic mode, for transparency (device-timed).  EVERY rank runs
    # this loop: device_step contains the all-reduce
    other = "reference_order" if args.mode == "decomposed" else "decomposed"
    eng.set_mode(hw._ffi.MODE_REFERENCE_ORDER if other == "reference_order" else hw._ffi.MODE_DECOMPOSED)
    for i in range(3):
        device_step(1000 + i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_other = min(args.steps, 50)
    e0.record()
    for i in range(n_other):
        device_step(5000 + i)
    e1.record()
    torch.cuda.synchronize()
    other_ms = e0.elapsed_time(e1) / n_other
    eng.set_mode(hw._ffi.MODE_DECOMPOSED if args.mode == "decomposed" else hw._ffi.MODE_REFERENCE_ORDER)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline denominators measured here: issue rate and pipe rates of this GPU ----
    def rate(which, iters=512):
        ms, n = eng.pipe_probe(which, iters)
        return n / (ms * 1e-3)
    ffma, ffma2, mufu, alu, i2f, mix = rate(0), rate(1), rate(2), rate(3), rate(4, 128), rate(6, 128)
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    sm_count = torch.cuda.get_device_properties(local_rank).multi_processor_count
    issue_peak_nominal = sm_count * 4 * 32 * sm_mhz * 1e6        # lane-instructions/s at the clock seen
    issue_peak = max(ffma, alu, mix)                              # measured: best single-issue stream
    per_gpu = value / world
    algo = ALGO[args.mode]
    xu_frac = per_gpu * algo["xu"] / mufu
    issue_frac = per_gpu * algo["issue"] / issue_peak
    bound = "xu" if xu_frac >= issue_frac else "issue"
    roofline = {
        # this path is bound by the SFU (XU) pipe and by instruction issue, not by HBM or tensor cores
        # (SURVEY 8d); "achieved"/"peak" are for the binding resource, the other one is given beside it
        "bound": bound,
        "achieved": (per_gpu * algo["xu"] if bound == "xu" else per_gpu * algo["issue"]) / 1e9,
        "peak": (mufu if bound == "xu" else issue_peak) / 1e9,
        "unit": "G MUFU/s (thread-level)" if bound == "xu" else "Ginstr/s (thread-level)",
        "frac": max(xu_frac, issue_frac),
        # dram__bytes_read.sum + dram__bytes_write.sum of one simulation-kernel launch
        # (profiles/r01_ncu_full_*.csv): window tables, L2-resident after first touch
        "traffic": 6635008, "traffic_unit": "bytes per launch (ncu --set full)",
        "peak_source": "measured by hw1f_pipe_probe on this GPU in this run (MUFU.EX2 / FFMA / LOP3 / mixed streams); "
                       "MEASURED_PEAKS.json has no FP32/XU entry (HBM and bf16 tensor only)",
        "mode": args.mode,
        "xu_pipe": {"achieved": per_gpu * algo["xu"] / 1e9, "peak_mufu": mufu / 1e9, "frac": xu_frac},
        "issue": {"achieved": per_gpu * algo["issue"] / 1e9, "peak_measured": issue_peak / 1e9,
                  "peak_nominal_at_clock": issue_peak_nominal / 1e9, "frac": issue_frac},
        "fp32_pipe": {"achieved": per_gpu * algo["fp32"] / 1e9, "peak_ffma": ffma / 1e9,
                      "peak_ffma2_lanes": 2 * ffma2 / 1e9, "frac": per_gpu * algo["fp32"] / max(ffma, 2 * ffma2)},
        "probes_Ginstr_s": {"ffma": ffma / 1e9, "ffma2": ffma2 / 1e9, "mufu_ex2": mufu / 1e9, "lop3_shf": alu / 1e9,
                            "i2fp": i2f / 1e9, "hw1f_mix": mix / 1e9},
        "algorithmic_per_path_step": algo,
        # static, from the committed ncu --set full capture of this kernel (not measured in this run): the dispatch
        # port is the binding resource -- packed FP32x2 instructions hold it for two cycles, so its utilisation is
        # issue-active + (fma-pipe cycles - fma instructions)
        "ncu_capture": ({"file": "profiles/r01_ncu_full_fast_kernel_v3.csv", "kernel_us": 572.7, "issue_active_pct": 81.2,
                         "dispatch_port_pct": 94.0, "xu_pipe_pct": 80.7, "alu_pipe_pct": 65.6, "fma_pipe_cycles_pct": 43.5,
                         "top_stall": "not_selected"} if args.mode == "decomposed" else
                        {"file": "profiles/r01_ncu_full_bond_curve_v3.csv"}),
        "kernel": ("fast_kernel<1,0,0>" if args.mode == "decomposed" else "bond_curve_kernel<1>") +
                  " (prep_lo_kernel + reduce_curve_kernel included in the time)",
        "other_mode": {"mode": other, "ms_per_step": other_ms,
                       "value": path_steps_per_step / world / (other_ms * 1e-3),
                       "issue_frac": path_steps_per_step / world / (other_ms * 1e-3) * ALGO[other]["issue"] / issue_peak},
    }

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle_lib import Oracle
        o = Oracle()
        # bounded sample of the same workload sized for ~10 s of wall time on this box's cores
        # (BASELINE.json configs[0] runs the oracle at 2^16; more cores -> a bigger sample)
        t0 = time.time()
        o.bond_curve(1, 1 << 14)
        t_cal = max(time.time() - t0, 1e-3)
        lg = 14
        while lg < 22 and t_cal * (1 << (lg + 1 - 14)) < 12.0:
            lg += 1
        n_cpu = 1 << max(lg, 16)
        t0 = time.time()
        o.bond_curve(1234, n_cpu)
        dt = time.time() - t0
        cpu_baseline = {"value": 2.0 * n_cpu * n_steps / dt, "unit": UNIT, "cores": o.max_threads(), "kind": "port",
                        "sample": f"OpenMP C oracle (oracle/hw1f_oracle.c), Q1 at 2^{n_cpu.bit_length() - 1} "
                                  f"subsequences x 2 x 1000 steps, {dt:.2f} s wall"}

    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clocks.get("reasons", []))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Q1 bond curve P(0,T), f(0,T): 2^{args.paths_log2} XORWOW subsequences x 2 antithetic "
                               f"paths x {n_steps} steps per GPU, r0=0.012 a=1 sigma=0.1, {n_mat} maturities, "
                               "seeding included", "arithmetic": args.mode,
                   "paths_per_gpu": 2 * n_paths, "n_steps": n_steps, "l2": "flushed between steps (256 MiB memset, "
                   "outside the per-step event pairs)", "parallelism": (f"path-range sharding x{world}, one all-reduce of 202 doubles per step "
                                                           f"({collective})") if world > 1 else "single GPU"},
        "collective": {"kind": collective, "check": collective_check},
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                # set_model uploads ONE arena: 2 duplicated drift tables + centring + exp(-Im), 256-byte aligned pieces
                "h2d_bytes_per_step": 2 * (((n_steps + 2) * 8 + 255) // 256 * 256) + 2 * ((n_mat * 4 + 255) // 256 * 256),
                "d2h_bytes_per_step": 3 * n_mat * 4},
        "gpu_launches": int(launches),
        "clocks": clocks, "clock_check": "rejected: thermal/hw slowdown seen" if bad else "ok",
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "check": {"P_0_10": float(last["P"][-1]), "f_0_0": float(last["f"][0])},
        "published_v100_path_steps_per_s": 3.91e11,
    }
    _emit(line, real_stdout)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
