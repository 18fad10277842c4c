#!/usr/bin/env bash
# source-level capture of the factorisation with the shortest sampling interval: what does the pivot-chain warp wait for?
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-pcg-block --no-e2e --no-phases --no-batch-block"
{
  timeout 600 ncu --set full --warp-sampling-interval 0 --warp-sampling-buffer-size 536870912 --clock-control none --import-source on -k "regex:k_potrf_df" -s 30 -c 1 -o gpurun_out/prof_chain $CMD > gpurun_out/ncu_chain.log 2>&1; tail -1 gpurun_out/ncu_chain.log | cut -c1-160
  ncu -i gpurun_out/prof_chain.ncu-rep --page source --csv --kernel-name regex:k_potrf_df > gpurun_out/potrf_chain_source.csv 2>/dev/null; wc -l gpurun_out/potrf_chain_source.csv
} > gpurun_out/round47.log 2>&1
cat gpurun_out/round47.log
