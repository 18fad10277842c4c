#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== timeline tt (panel hand-off)"; timeout 120 scripts/bin/df_timeline_tt 1000 > gpurun_out/timeline_tt.log; grep -E "^rep|^info|step [0-3]|last tile" gpurun_out/timeline_tt.log; grep -A2 "^chain   [3-5] " gpurun_out/timeline_tt.log
  for n in 500 1000 1280 4096; do echo "== potrf la1 n=$n"; timeout 120 scripts/bin/df_timeline_la1 $n | grep -E "^rep|^info"; done
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  for s in 16 32; do
    echo "== bnb slots $s windows"; timeout 300 python bench.py --workload bnb --slots $s --steps 20 --warmup 3 2>> gpurun_out/bnb.err | tee gpurun_out/bnb_s$s.json | cut -c1-150
    echo "== bnb slots $s stream x4"; timeout 300 python bench.py --workload bnb --slots $s --steps 5 --warmup 3 --stream-factor 4 2>> gpurun_out/bnb.err | tee gpurun_out/bnb_stream_s$s.json | cut -c1-150
  done
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-solve 2> gpurun_out/bench_o.err | tee gpurun_out/bench_o.json | cut -c1-300
} > gpurun_out/round24.log 2>&1
cat gpurun_out/round24.log
