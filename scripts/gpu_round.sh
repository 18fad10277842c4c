#!/usr/bin/env bash
# bench line + ncu launch list + ncu --set full on the three dominant direct-path kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pcg-block --no-e2e --no-phases"
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err
$CMD > gpurun_out/plain_c.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c.csv $CMD > gpurun_out/ncu_c1.log 2>&1
$CMD > gpurun_out/plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_potrf_df|k_trsv_df|k_assemble_normal" -s 30 -c 6 -o gpurun_out/prof_direct $CMD > gpurun_out/ncu_c2.log 2>&1
cat gpurun_out/bench_f.json; tail -3 gpurun_out/ncu_c1.log gpurun_out/ncu_c2.log
