#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== pytest gpu bnb"; timeout 900 python -m pytest tests/test_gpu_bnb.py -x -q -m gpu 2>&1 | tail -3
  echo "== bnb slots 32 windows + stream"; for f in 0 4; do timeout 300 python bench.py --workload bnb --slots 32 --steps $((20 - 3*f)) --warmup 3 --stream-factor $f 2>> gpurun_out/bnb.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['nodes'], d['incumbent'], d['root_bound'])"; done
} > gpurun_out/round46.log 2>&1
cat gpurun_out/round46.log
