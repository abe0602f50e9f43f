#!/bin/bash
# One GPU-box session: tests, bench (both arms), launch list and ncu --set full captures.  Usage: tools/gpu_round.sh [tag]
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
echo "== pytest gpu" ; python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
echo "== bench reference" ; python bench.py --impl reference --steps 20 --warmup 3 | tee $OUT/${TAG}_bench_ref.json | cut -c1-300
echo "== bench" ; python bench.py 2> $OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-400
echo "== api overhead" ; python tools/api_overhead.py > $OUT/${TAG}_api_overhead.json 2>&1
echo "== fixed cost" ; python tools/fixed_cost_probe.py 2>&1 | tee $OUT/${TAG}_fixed_cost.txt
echo "== submission lanes" ; python tools/two_engine_probe.py 300 2>&1 | tee $OUT/${TAG}_submit_collect_lanes.txt
echo "== launch list (only after the plain run above exited)"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches_q1_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-workloads --no-scaling-run --no-cpu-baseline > /dev/null 2>&1
echo "== ncu --set full: Q1 simulation kernel, tail kernel, prep_lo kernel"
ncu --set full --clock-control none --import-source on -k regex:fast_kernel -s 4 -c 1 -o $OUT/${TAG}_ncu_fast_kernel -f \
    python bench.py --steps 3 --warmup 3 --no-workloads --no-scaling-run --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none -k regex:tail_kernel -s 4 -c 1 -o $OUT/${TAG}_ncu_tail_kernel -f \
    python bench.py --steps 3 --warmup 3 --no-workloads --no-scaling-run --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none -k regex:prep_lo_kernel -s 4 -c 1 -o $OUT/${TAG}_ncu_prep_lo_kernel -f \
    python bench.py --steps 3 --warmup 3 --no-workloads --no-scaling-run --no-cpu-baseline > /dev/null 2>&1
echo "== ncu --set full: Q3 one-launch sequence kernel, ZBC kernel"
ncu --set full --clock-control none -k regex:fast_kernel -s 2 -c 1 -o $OUT/${TAG}_ncu_q3_sequence -f python tools/run_q3.py > /dev/null 2>&1
ncu --set full --clock-control none -k regex:fast_kernel -s 1 -c 1 -o $OUT/${TAG}_ncu_fast_zbc -f python tools/run_zbc_once.py zbc > /dev/null 2>&1
ls -la $OUT/${TAG}_ncu_*.ncu-rep
