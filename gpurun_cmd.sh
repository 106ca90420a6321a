python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_r01_l.json 2>/dev/null;  python -c "
import json
r=json.load(open('gpurun_out/bench_r01_l.json')); print('1GPU VALUE %.4e cu/s'%r['value'], r['ms_per_step'], r['phase_ms'], r['clocks'])"
