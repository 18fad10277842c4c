#!/usr/bin/env bash
mkdir -p gpurun_out
{
  for v in tt tt0; do echo "== tile timing $v"; timeout 120 scripts/bin/df_timeline_$v 1000 | grep -E "^rep|^info|step [0-3]|panel [0-3]|last tile"; done
  for n in 1000 1280; do echo "== potrf la1 n=$n"; timeout 120 scripts/bin/df_timeline_la1 $n | grep -E "^rep|^info"; done
  echo "== potrf noz n=1280"; timeout 120 scripts/bin/df_timeline_noz 1280 | grep -E "^rep|^info"
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  for s in 8 16 32; do
    echo "== bnb slots $s device"; timeout 300 python bench.py --workload bnb --slots $s --steps 20 --warmup 3 2>> gpurun_out/bnb.err | tee gpurun_out/bnb_s$s.json | cut -c1-150
  done
  for s in 16 32; do
    echo "== bnb slots $s stream x4"; timeout 300 python bench.py --workload bnb --slots $s --steps 5 --warmup 3 --stream-factor 4 2>> gpurun_out/bnb.err | tee gpurun_out/bnb_stream_s$s.json | cut -c1-150
  done
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-solve 2> gpurun_out/bench_o.err | tee gpurun_out/bench_o.json | cut -c1-300
} > gpurun_out/round23.log 2>&1
cat gpurun_out/round23.log
