#!/usr/bin/env bash
mkdir -p gpurun_out
{
  for v in 0 1; do for n in 1000 4096; do echo "== potrf lookahead=$v n=$n"; timeout 120 scripts/bin/df_timeline_la$v $n | grep -E "^rep|^info"; done; done
  echo "== timeline lookahead=1"; timeout 120 scripts/bin/df_timeline_la1 1000 > gpurun_out/timeline_la1.log; grep -A3 "^chain   [345]" gpurun_out/timeline_la1.log
  echo "== pytest gpu kernels+bnb"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_bnb.py -x -q -m gpu 2>&1 | tail -15
  for s in 8 16 32 64; do
    echo "== bnb slots $s device"; timeout 300 python bench.py --workload bnb --slots $s --steps 20 --warmup 3 2>> gpurun_out/bnb.err | cut -c1-400
  done
  echo "== bnb slots 16 host"; timeout 300 python bench.py --workload bnb --slots 16 --steps 20 --warmup 3 --host-heuristics 2>> gpurun_out/bnb.err | cut -c1-400
  echo "== pytest gpu all"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-solve 2> gpurun_out/bench_o.err | tee gpurun_out/bench_o.json | cut -c1-300
} > gpurun_out/round22.log 2>&1
cat gpurun_out/round22.log
