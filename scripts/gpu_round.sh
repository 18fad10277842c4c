#!/usr/bin/env bash
mkdir -p gpurun_out
{
  for v in la1 inl; do for n in 1000; do echo "== potrf $v n=$n"; timeout 120 scripts/bin/df_timeline_$v $n | grep -E "^rep|^info"; done; done
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-solve 2> gpurun_out/bench_o.err | tee gpurun_out/bench_o.json | cut -c1-300
  for w in scp4x scpnrf scpclr13; do echo "== bench $w"; timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-pcg-block 2>> gpurun_out/bench_o.err | tee gpurun_out/bench_$w.json | cut -c1-200; done
} > gpurun_out/round26.log 2>&1
cat gpurun_out/round26.log
