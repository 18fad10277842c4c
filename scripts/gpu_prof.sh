#!/usr/bin/env bash
# ncu evidence pass: launch list of the default bench line (shares of the step) and --set full captures of the
# one-block solver and the reference-rules node kernel.   bash scripts/gpu_prof.sh <tag>
tag=${1:-prof}
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-pcg-block --no-bnb-block --no-batch-block --no-e2e"
$B > gpurun_out/${tag}_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches_bench.csv $B > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "bench launch list rc=$?"
P="python scripts/prof_cta.py scpnrh1 2"
$P > gpurun_out/${tag}_plain_cta.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_ipm_cta|k_node_heuristics_ref" -c 4 -o gpurun_out/${tag}_cta $P > gpurun_out/${tag}_ncu_cta.log 2>&1
echo "cta capture rc=$?"; cat gpurun_out/${tag}_plain_cta.log; ls -la gpurun_out/ | tail -8
