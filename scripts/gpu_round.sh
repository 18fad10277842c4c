#!/usr/bin/env bash
# One GPU-box round: full GPU test-suite and the bench line.  Logs -> gpurun_out/.
mkdir -p gpurun_out
{
  echo "== timeline"; timeout 60 scripts/bin/df_timeline 1024 | grep "^rep\|residual"
  echo "== timeline 4096"; timeout 60 scripts/bin/df_timeline 4096 | grep "^rep\|residual"
  echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
  echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/bench_e.err | tee gpurun_out/bench_e.json
} > gpurun_out/round7.log 2>&1
tail -c 5000 gpurun_out/round7.log
