#!/usr/bin/env bash
# GPU-box check of the reference builds (oracle/_ref): the reference's own CUDA solver and the drop-in build
# of the same tree, on instance files written from the committed fixtures.
mkdir -p gpurun_out /tmp/sb200_data
python - <<'PY'
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_golden
from oracle import scp_io
for nm in ("scp41", "scp48", "scp410", "scpnre1", "scpnrh1", "scpclr10"):
    inst, _ = load_golden(nm)
    scp_io.write_scp_text(inst, f"/tmp/sb200_data/{nm}.txt")
PY
R=oracle/_ref
{
  echo "== fp64 peak"; timeout 120 scripts/bin/fp64_peak
  for nm in scp41 scpclr10; do
  for v in ref b200; do
    echo "== sypha_$v $nm LP"; timeout 300 $R/sypha_$v --model scp --input-file /tmp/sb200_data/$nm.txt --mehrotra-max-iter 100 --disable-bnb --verbosity 5 2>&1 | tail -8
    echo "== api_lp_$v $nm --lp"; timeout 300 $R/api_lp_$v /tmp/sb200_data/$nm.txt --lp --max-iter 100 2>&1 | tail -2
  done; done
  echo "== sypha_ref scpnrh1 5 iterations"; timeout 600 $R/sypha_ref --model scp --input-file /tmp/sb200_data/scpnrh1.txt --mehrotra-max-iter 5 --disable-bnb --verbosity 5 2>&1 | tail -8
  echo "== sypha_b200 scpnrh1"; timeout 600 $R/sypha_b200 --model scp --input-file /tmp/sb200_data/scpnrh1.txt --mehrotra-max-iter 100 --disable-bnb --verbosity 5 2>&1 | tail -8
  echo "== sypha_ref scpnre1 100 iterations"; timeout 600 $R/sypha_ref --model scp --input-file /tmp/sb200_data/scpnre1.txt --mehrotra-max-iter 100 --disable-bnb --verbosity 5 2>&1 | tail -8
  for v in b200 ref; do
    echo "== api_lp_$v scp48 B&B"; timeout 400 $R/api_lp_$v /tmp/sb200_data/scp48.txt --max-iter 100 --time-limit 120 --verbosity 5 2>&1 | tail -12
  done
  echo "== api_lp_b200 scp410 B&B"; timeout 400 $R/api_lp_b200 /tmp/sb200_data/scp410.txt --max-iter 100 --time-limit 120 --verbosity 5 2>&1 | tail -6
} > gpurun_out/ref_check.log 2>&1
echo "== pytest gpu" >> gpurun_out/ref_check.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 >> gpurun_out/ref_check.log
cat gpurun_out/ref_check.log
