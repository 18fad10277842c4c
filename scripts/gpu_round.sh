#!/usr/bin/env bash
# final evidence pass: bench line, ncu launch list, ncu --set full of the dominant kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pcg-block --no-e2e --no-phases"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
$CMD > gpurun_out/plain_f.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_f.csv $CMD > gpurun_out/ncu_f1.log 2>&1
$CMD > gpurun_out/plain_f2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_potrf_df|k_trsv_df|k_assemble_normal16_smem|k_spmv_csc|k_spmv_csr" -s 40 -c 8 -o gpurun_out/prof_final $CMD > gpurun_out/ncu_f2.log 2>&1
python scripts/time_pcg_phases.py --reps 1 > gpurun_out/plain_f3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_blk_rows|k_blk_cols|k_cg_" -s 6 -c 6 -o gpurun_out/prof_final_pcg python scripts/time_pcg_phases.py --reps 1 > gpurun_out/ncu_f3.log 2>&1
cat gpurun_out/bench_final.json | cut -c1-600
