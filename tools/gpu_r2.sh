#!/bin/bash
# Round-2 GPU passes (run under gpurun from the repo root):  tools/gpu_r2.sh <stage> [tag]
#   tests     GPU parity suite + a few extra seeds of the random-table property test
#   bench     the contract line (all BASELINE configs inside) + the reference arm
#   sweep     every golden table, random / sorted (/ clustered for config 4) points
#   ab        A/B switches of this round's layout choices on the tables they target
#   ncu NAME WORKLOAD [extra bench args]   one `--set full` capture of the eval kernel; the raw and
#             source pages are exported as CSV on the box (the .ncu-rep stays there: 64 MiB pull cap)
#   launches  the ncu launch list of the contract command
# Every ncu pass runs only after the same command exited 0 without ncu.
set -x
stage=$1
tag=${2:-r2}
mkdir -p gpurun_out
case "$stage" in
tests)
    python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; tail -5 gpurun_out/pytest_${tag}.log
    for sd in 1 2 3; do AI_TEST_SEED=$sd python -m pytest tests -m gpu -x -q -k lookup_exact_on_random_tables 2>&1 | tail -1; done
    ;;
bench)
    python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err || { tail -20 gpurun_out/bench_${tag}.err; exit 1; }
    python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2>&1
    ;;
sweep)
    python tools/sweep.py --tables all --tag sweep_${tag} > gpurun_out/sweep_${tag}.txt 2>&1
    tail -130 gpurun_out/sweep_${tag}.txt
    ;;
ab)
    T=approximations__64o5-p2.22044604925e-14,gen__cfg4_f1_leg_o7_f64,mx__cheb_o3_f64,mx__leg_o3_f64
    python tools/sweep.py --tables $T --tag ab_vec_on_${tag} > gpurun_out/ab_vec_on_${tag}.txt 2>&1
    AI_B200_NO_VEC_MODE=1 python tools/sweep.py --tables $T --tag ab_vec_off_${tag} > gpurun_out/ab_vec_off_${tag}.txt 2>&1
    T=approx__32o3-p1.19209289551e-06,approximations__32o3-p1.19209289551e-06,approximations__32o3-p1.19209289551e-05,mx__cheb_o3_f32,mx__cheb_o4_f32,approximations__64o5-p2.22044604925e-14
    python tools/sweep.py --tables $T --tag ab_rank16_on_${tag} > gpurun_out/ab_rank16_on_${tag}.txt 2>&1
    AI_B200_RANK16_MIN_BYTES=1000000 python tools/sweep.py --tables $T --tag ab_rank16_off_${tag} > gpurun_out/ab_rank16_off_${tag}.txt 2>&1
    tail -8 gpurun_out/ab_*_${tag}.txt
    ;;
compact)
    T=large__approx__64o3-p2.22044604925e-14,large__approx__64o3-p2.22044604925e-15,large__approximations__64o3-p2.22044604925e-14,synth:16384:3:f64,synth:131072:3:f64,synth:16384:7:f64,synth:16384:3:f32
    python tools/sweep.py --tables $T --tag compact_on_${tag} > gpurun_out/compact_on_${tag}.txt 2>&1
    AI_B200_NO_COMPACT=1 python tools/sweep.py --tables $T --tag compact_off_${tag} > gpurun_out/compact_off_${tag}.txt 2>&1
    grep -v "^copy\|^table" gpurun_out/compact_on_${tag}.txt gpurun_out/compact_off_${tag}.txt
    ;;
ncu)
    name=$3; workload=$4; shift 4
    args="--workload $workload --steps 4 --warmup 3 --no-e2e --no-cpu --no-configs --no-sustained $*"
    python bench.py $args > gpurun_out/plain_${name}.log 2>&1 || { tail -5 gpurun_out/plain_${name}.log; exit 1; }
    ncu --set full --clock-control none --import-source on -k regex:eval_ -s 4 -c 1 -f -o /tmp/prof_${name} python bench.py $args > gpurun_out/ncu_${name}.log 2>&1
    ncu -i /tmp/prof_${name}.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_${name}_raw.csv
    ncu -i /tmp/prof_${name}.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_${tag}_${name}_source.csv.gz
    ls -la /tmp/prof_${name}.ncu-rep gpurun_out/ncu_${tag}_${name}_*
    ;;
multi)
    python tools/multi_bench.py --tag multi_${tag} > gpurun_out/multi_${tag}.txt 2>&1; cat gpurun_out/multi_${tag}.txt
    ;;
launches)
    python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/plain_launches.log 2>&1 || exit 1
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/ncu_l.log 2>&1
    ;;
*)
    echo "unknown stage $stage"; exit 2
    ;;
esac
