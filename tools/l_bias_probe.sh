for b in ${BIASES:-0 1 2 3}; do
echo "== HW1F_L_BIAS=$b"
HW1F_L_BIAS=$b python tools/fixed_cost_probe.py 2>&1 | tail -4 | tr '\n' ' '; echo
HW1F_L_BIAS=$b python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-workloads --no-scaling-run 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); print('ms/step %.4f' % l['ms_per_step'], 'e2e_ms %.4f' % l['e2e']['ms_per_step'], 'sustained %.4f' % l['sustained']['ms_per_step'], 'fixed %.4f' % l['roofline']['steady_state']['fixed_ms_per_call'])"
done
