set -u
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r04c_pytest_multi.txt
port=29600
for n in 8 4 2; do
  port=$((port+1)); $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --steps 100 --warmup 5 2>gpurun_out/r04c_bench$n.err > gpurun_out/r04c_bench$n.json
  port=$((port+1)); $TR --nproc-per-node $n --master-port $port tools/scaling_run.py --total-log2 30 2>gpurun_out/r04c_scaling$n.err > gpurun_out/r04c_scaling_${n}gpu.json
done
python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>gpurun_out/r04c_bench1.err > gpurun_out/r04c_bench1.json
python tools/scaling_run.py --total-log2 30 2>gpurun_out/r04c_scaling1.err > gpurun_out/r04c_scaling_1gpu.json
python - <<'PY'
import json
for n in (1,2,4,8):
    b=json.load(open(f"gpurun_out/r04c_bench{n}.json")); s=json.load(open(f"gpurun_out/r04c_scaling_{n}gpu.json"))
    print(n, round(b["ms_per_step"],4), "%.4e"%b["value"], "%.4e"%b["e2e"]["value"], "|", round(s["seconds"],5), "%.4e"%s["path_steps_per_s"], s["P_0_10"], s["zbc_price_cv"], s["vega_pathwise"])
PY
