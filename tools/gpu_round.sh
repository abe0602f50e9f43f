#!/bin/bash
# One GPU-box session: tests, reference probes, pipe probes, bench.  Usage: tools/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $OUT/${TAG}_gpu.txt 2>&1
nproc >> $OUT/${TAG}_gpu.txt
echo "== probe_steps" ; oracle/_ref/probe_steps | tee $OUT/${TAG}_probe_steps.jsonl
echo "== pytest gpu" ; python -m pytest tests -x -q -m gpu 2>&1 | tail -25 | tee $OUT/${TAG}_pytest.txt
echo "== pipe probe" ; python tools/pipe_probe.py | tee $OUT/${TAG}_pipe_probe.json
echo "== ref harness bench"
for q in q1 q2 q3; do oracle/_ref/ref_harness bench $q 20 3 $OUT/${TAG}_ref_${q}.json > /dev/null 2>&1; cat $OUT/${TAG}_ref_${q}.json; done
echo "== bench" ; python bench.py --steps 100 --warmup 5 | tee $OUT/${TAG}_bench.json
echo "== bench reference" ; python bench.py --impl reference --steps 20 --warmup 3 | tee $OUT/${TAG}_bench_ref.json
