#!/usr/bin/env bash
mkdir -p gpurun_out
{
  for v in la1 piv1; do for n in 1000 300; do echo "== potrf $v n=$n"; timeout 120 scripts/bin/df_timeline_$v $n | grep -E "^rep|^info"; done; done
  echo "== tile timing (piv2)"; timeout 120 scripts/bin/df_timeline_tt 1000 | grep -E "step [0-3]|last tile"
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-block 2> gpurun_out/bench_o.err | tee gpurun_out/bench_o.json | cut -c1-300
} > gpurun_out/round28.log 2>&1
cat gpurun_out/round28.log
