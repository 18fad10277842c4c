#!/usr/bin/env bash
mkdir -p gpurun_out
{
  echo "== pytest blocked + pcg"; timeout 900 python -m pytest tests -x -q -m gpu -k "blocked or general_coeff or pcg or spmv" 2>&1 | tail -15
  echo "== synth50k blocked"; timeout 300 python scripts/run_synth.py --max-iter 3 2>&1 | tail -6
  echo "== synth50k value kernels"; SB200_BLOCKED=0 timeout 300 python scripts/run_synth.py --max-iter 3 2>&1 | tail -6
} > gpurun_out/round8.log 2>&1
cat gpurun_out/round8.log
