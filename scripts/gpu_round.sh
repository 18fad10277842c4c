#!/usr/bin/env bash
mkdir -p gpurun_out
{
  for v in la1 w4 la1 w4; do echo "== potrf $v (la1: warp 4 idle, w4: warp 4 works) n=1000"; timeout 120 scripts/bin/df_timeline_$v 1000 | grep -E "^rep|^info"; done
  echo "== pytest gpu kernels+solve"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py -x -q -m gpu 2>&1 | tail -3
  echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-pcg-block --no-batch-block 2> gpurun_out/bench_o.err | tee gpurun_out/bench_o.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['loop_ms_per_lp'],3), {k:round(v['ms']*1e3,1) for k,v in d['phases'].items()})"
} > gpurun_out/round43.log 2>&1
cat gpurun_out/round43.log
