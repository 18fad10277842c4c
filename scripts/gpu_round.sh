#!/usr/bin/env bash
# final evidence pass: bench line, ncu launch list, ncu --set full of the dominant kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pcg-block --no-e2e --no-phases"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err
$CMD > gpurun_out/plain_f.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_g.csv $CMD > gpurun_out/ncu_g1.log 2>&1
$CMD > gpurun_out/plain_f2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:k_potrf_df|k_trsv_df|k_assemble_normal16_smem|k_spmv_csc|k_spmv_csr|k_update" -s 40 -c 10 -o gpurun_out/prof_final2 $CMD > gpurun_out/ncu_g2.log 2>&1
cat gpurun_out/bench_final2.json | cut -c1-300
