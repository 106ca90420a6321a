python -m pytest tests -m gpu -x -q 2>&1 | tail -12
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_r01_f.json 2> gpurun_out/bench_r01_f.err; python -c "
import json
r=json.load(open('gpurun_out/bench_r01_f.json')); print('VALUE %.3e cu/s'%r['value'], r['phase_ms'], r['roofline']['whole_call'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweby -s 9 -c 3 -o gpurun_out/prof_r01_f python bench.py --case global_025deg --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof.log 2>&1; tail -1 gpurun_out/prof.log
