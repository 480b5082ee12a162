set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; tail -c 600 gpurun_out/bench_r1k.json
python bench.py --workload cfg3_j0_cheb_o13_f64 --steps 50 --no-e2e --no-cpu > gpurun_out/bench_r1k_cfg3.json 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1k_ref.json 2>&1
# launch list (only after the plain command above exited 0)
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1k.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/ncu_l.log 2>&1
# full captures of the eval kernel
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 1 -f -o gpurun_out/prof_r1k_cfg2 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 1 -f -o gpurun_out/prof_r1k_cfg3 python bench.py --workload cfg3_j0_cheb_o13_f64 --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
