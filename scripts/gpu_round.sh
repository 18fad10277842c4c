#!/usr/bin/env bash
mkdir -p gpurun_out
rm -f gpurun_out/bench_ladder.jsonl
timeout 600 python bench.py --workload synth5k --strategy cholesky --steps 3 --warmup 3 --no-pcg-block --no-cpu-baseline --no-phases --no-e2e 2>> gpurun_out/bench_ladder.err >> gpurun_out/bench_ladder.jsonl
for tol in 1e-8 1e-6; do
timeout 900 python bench.py --workload synth5k --strategy pcg --cg-tol $tol --steps 3 --warmup 3 --no-pcg-block --no-cpu-baseline --no-phases --no-e2e 2>> gpurun_out/bench_ladder.err >> gpurun_out/bench_ladder.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_ladder.jsonl'):
    d=json.loads(l); print(d['config']['strategy'], round(d['value'],1), 'iter/s', round(d['time_to_lp_opt_ms'],1), 'ms/LP', round(d['iterations_per_lp'],2), 'it/LP', 'launches', d['gpu_launches'])
PY
timeout 300 python scripts/run_synth.py --max-iter 100 --cg-max-iter 200000 --cg-tol 1e-6 2>&1 | tail -2
