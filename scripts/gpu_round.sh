#!/usr/bin/env bash
# evidence pass: full bench line, ncu launch list, ncu --set full of the dominant kernels, bnb lines
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pcg-block --no-e2e --no-phases --no-batch-block"
{
  echo "== bench full"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; cut -c1-300 gpurun_out/bench_k.json
  echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_k_ref.json 2>> gpurun_out/bench_k.err; cut -c1-300 gpurun_out/bench_k_ref.json
  echo "== launch list"; $CMD > gpurun_out/plain_k.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_k.csv $CMD > gpurun_out/ncu_k1.log 2>&1; tail -2 gpurun_out/ncu_k1.log | cut -c1-200
  echo "== ncu full"; $CMD > gpurun_out/plain_k2.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:k_potrf_df|k_tri_gemv|k_assemble_normal16_smem|k_spmv_csc|k_spmv_csr|k_update" -s 40 -c 12 -o gpurun_out/prof_k $CMD > gpurun_out/ncu_k2.log 2>&1; tail -2 gpurun_out/ncu_k2.log | cut -c1-200
  python scripts/ncu_summary.py gpurun_out/prof_k.ncu-rep > gpurun_out/ncu_k_summary.txt 2>&1; head -30 gpurun_out/ncu_k_summary.txt
  for s in 16 32; do
    echo "== bnb slots $s windows"; timeout 300 python bench.py --workload bnb --slots $s --steps 20 --warmup 3 2>> gpurun_out/bnb.err | tee gpurun_out/bnb_k_s$s.json | cut -c1-150
    echo "== bnb slots $s stream x4"; timeout 300 python bench.py --workload bnb --slots $s --steps 5 --warmup 3 --stream-factor 4 2>> gpurun_out/bnb.err | tee gpurun_out/bnb_k_stream_s$s.json | cut -c1-150
  done
  echo "== bnb heuristics kernel under ncu"; timeout 300 ncu --set full --clock-control none -k "regex:k_node_heuristics" -s 20 -c 3 -o gpurun_out/prof_k_heur python bench.py --workload bnb --slots 8 --steps 6 --warmup 1 > gpurun_out/ncu_k3.log 2>&1; python scripts/ncu_summary.py gpurun_out/prof_k_heur.ncu-rep > gpurun_out/ncu_k_heur_summary.txt 2>&1; head -12 gpurun_out/ncu_k_heur_summary.txt
} > gpurun_out/round27.log 2>&1
cat gpurun_out/round27.log
