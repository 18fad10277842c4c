#!/usr/bin/env bash
# launch list of the B&B workload (shares of a round): python bench.py --workload bnb ...
mkdir -p gpurun_out
B="python bench.py --workload bnb --bnb-instance scpnrg1 --steps 4 --warmup 2"
$B > gpurun_out/bnbprof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/bnbprof_launches.csv $B > gpurun_out/bnbprof_ncu.log 2>&1
echo "rc=$?"; tail -c 600 gpurun_out/bnbprof_plain.log; grep -c '^"' gpurun_out/bnbprof_launches.csv
