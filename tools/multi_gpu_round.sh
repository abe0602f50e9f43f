#!/bin/bash
# On an 8-GPU box: the driver's scaling shape (bench.py at N = 8, 4, 2, 1; every line carries the 2^30 configs[4]
# scaling_run and, for N > 1, the multi-vs-single-GPU moment check) + the multi-GPU tests.  Usage: tools/multi_gpu_round.sh [tag]
set -u
TAG=${1:-r02}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest_multi.txt
port=29600
for n in 8 4 2; do
  port=$((port+1)); timeout 600 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 100 --warmup 5 2>gpurun_out/${TAG}_bench$n.err > gpurun_out/${TAG}_bench_${n}gpu.json; echo "N=$n rc=$?"
done
python bench.py --steps 100 --warmup 5 --no-workloads --no-cpu-baseline 2>gpurun_out/${TAG}_bench1.err > gpurun_out/${TAG}_bench_1gpu_same_box.json; echo "N=1 rc=$?"
python - <<PY
import json
for n, name in ((1, "1gpu_same_box"), (2, "2gpu"), (4, "4gpu"), (8, "8gpu")):
    b = json.load(open(f"gpurun_out/${TAG}_bench_{name}.json")); s = b["scaling_run"]
    print(n, "ms/step %.4f" % b["ms_per_step"], "value %.4e" % b["value"], "e2e %.4e" % b["e2e"]["value"], "| 2^30:", "%.5f s" % s["seconds"],
          "%.4e" % s["path_steps_per_s"], s["P_0_10"], s["zbc_price_cv"], s["vega_pathwise"], "| check", b["check"]["multi_vs_single_max_rel"],
          b["collective"])
PY
