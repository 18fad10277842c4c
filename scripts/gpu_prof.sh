#!/usr/bin/env bash
# ncu evidence pass: launch list of the default bench line (shares of the step) and a --set full capture of one whole
# window of the throughput headline (148 thread blocks of k_ipm_cta).   bash scripts/gpu_prof.sh <tag>
tag=${1:-prof}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-pcg-block --no-bnb-block --no-e2e --no-phases --quick-single"
$B > gpurun_out/${tag}_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/${tag}_launches_bench.csv $B > gpurun_out/${tag}_ncu_bench.log 2>&1
echo "bench launch list rc=$?"
P="python scripts/prof_window.py 148 2"
$P > gpurun_out/${tag}_plain_window.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_ipm_cta" -s 1 -c 1 -o gpurun_out/${tag}_window $P > gpurun_out/${tag}_ncu_window.log 2>&1
echo "window capture rc=$?"; cat gpurun_out/${tag}_plain_window.log; tail -3 gpurun_out/${tag}_ncu_window.log; ls -la gpurun_out/ | tail -6
