#!/bin/bash
# One GPU pass that regenerates the evidence under profiles/ (run under gpurun):
# (gpurun copies back at most 64 MiB: the four --set full captures are split over two calls,
# `gpu_round_check.sh` and `gpu_round_check.sh ncu2`.)
# tests, contract bench lines, reference arm, all-table sweep, multi-table bench, then -- only after
# the plain commands exited 0 -- the ncu launch list and the --set full captures.
set -x
if [ "$1" = "ncu2" ]; then
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 1 -f -o gpurun_out/prof_final_cfg2_sorted python bench.py --order sorted --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 1 -f -o gpurun_out/prof_final_cfg4 python bench.py --workload cfg4_f1_leg_o7_f32 --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/
exit 0
fi
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for sd in 1 2 3 4 5; do AI_TEST_SEED=$sd python -m pytest tests -m gpu -x -q -k lookup_exact_on_random_tables 2>&1 | tail -1; done
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err || exit 1
python bench.py --order sorted --no-e2e --no-cpu --steps 50 > gpurun_out/bench_final_sorted.json 2>&1
python bench.py --workload cfg3_j0_cheb_o13_f64 --steps 50 --no-e2e --no-cpu > gpurun_out/bench_final_cfg3.json 2>&1
python bench.py --workload cfg4_f1_leg_o7_f32 --steps 50 --no-e2e --no-cpu > gpurun_out/bench_final_cfg4.json 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2>&1
python tools/sweep.py --tables all --tag sweep_final > gpurun_out/sweep_final.txt 2>&1
python tools/multi_bench.py --tag multi_final > gpurun_out/multi_final.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 1 -f -o gpurun_out/prof_final_cfg2 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 4 -c 1 -f -o gpurun_out/prof_final_cfg3 python bench.py --workload cfg3_j0_cheb_o13_f64 --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/
