#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
  for rep in 1 2; do
    echo "== bench main lib ($rep)"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-block --no-batch-block 2> gpurun_out/bench_o.err | tee gpurun_out/bench_main$rep.json | cut -c1-120
    echo "== bench alt lib: single-CTA vector kernels up to n = 16384 ($rep)"; SYPHA_B200_LIB=$PWD/sypha_b200/lib/libsypha_b200_alt.so timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-block --no-batch-block 2>> gpurun_out/bench_o.err | tee gpurun_out/bench_alt$rep.json | cut -c1-120
  done
  python - <<'PY'
import json
for f in ['bench_main1','bench_alt1','bench_main2','bench_alt2']:
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, round(d['value'],1), round(d['e2e']['value'],1), d['iterations_per_lp'], {k:round(v['ms']*1e3,1) for k,v in d['phases'].items()})
PY
} > gpurun_out/round29.log 2>&1
cat gpurun_out/round29.log
