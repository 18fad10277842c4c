#!/usr/bin/env bash
# refresh of the ncu --set full captures for the two kernels that changed after the r1_k evidence pass
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pcg-block --no-e2e --no-phases --no-batch-block"
{
  $CMD > gpurun_out/plain_m.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_potrf_df|k_tri_gemv" -s 20 -c 3 -o gpurun_out/prof_m $CMD > gpurun_out/ncu_m.log 2>&1; tail -1 gpurun_out/ncu_m.log | cut -c1-160
  python scripts/ncu_summary.py gpurun_out/prof_m.ncu-rep > gpurun_out/ncu_m_summary.txt 2>&1
  timeout 300 ncu --set full --clock-control none -k "regex:k_node_heuristics" -s 20 -c 3 -o gpurun_out/prof_m_heur python bench.py --workload bnb --slots 8 --steps 6 --warmup 1 > gpurun_out/ncu_m3.log 2>&1; python scripts/ncu_summary.py gpurun_out/prof_m_heur.ncu-rep >> gpurun_out/ncu_m_summary.txt 2>&1
  grep -E "Kernel Name|gpu__time_duration|dram__bytes_read|pipe_tensor|stall" gpurun_out/ncu_m_summary.txt | cut -c1-150
} > gpurun_out/round45.log 2>&1
cat gpurun_out/round45.log
