set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01_a.json 2> gpurun_out/bench_r01_a.err; tail -c 3000 gpurun_out/bench_r01_a.json; tail -5 gpurun_out/bench_r01_a.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_a.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_ncu.log 2>&1; tail -3 gpurun_out/bench_ncu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweby -s 9 -c 3 -o gpurun_out/prof_r01_a python bench.py --case global_025deg --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof.log 2>&1; tail -3 gpurun_out/prof.log
ls -la gpurun_out
