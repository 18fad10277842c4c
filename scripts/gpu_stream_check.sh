#!/usr/bin/env bash
mkdir -p gpurun_out
python - <<'PY'
import sys
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_golden
from oracle import scp_io
for nm in ("scpnre1", "scpnrg1"):
    inst, _ = load_golden(nm)
    scp_io.write_scp_text(inst, f"/tmp/{nm}.txt")
PY
{
for conn in 8 32; do
  export CUDA_DEVICE_MAX_CONNECTIONS=$conn
  for extra in "--slots 128" "--slots 128 --stream-factor 4" "--slots 64 --stream-factor 4"; do
    echo "== connections $conn: bnb scpnre1 $extra"; timeout 600 python bench.py --workload bnb --bnb-instance scpnre1 --steps 8 --warmup 3 $extra 2>>gpurun_out/stream_check.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['bnb']
print({a:(round(b[a],2) if isinstance(b[a],float) else b[a]) for a in ('value','nodes','lp_device_ms_per_node','ms_per_round')})"
  done
  for extra in "" "--continuous"; do
    echo "== connections $conn: C++ driver scpnre1 $extra"; timeout 300 oracle/_ref/bnb_batched_b200 /tmp/scpnre1.txt --max-iter 100 --max-nodes 3000 --slots 128 --no-preprocessing $extra 2>&1 | tail -1 | cut -c1-330
  done
done
echo "== C++ driver scpnrg1 windows"; timeout 300 oracle/_ref/bnb_batched_b200 /tmp/scpnrg1.txt --max-iter 100 --max-nodes 2000 --slots 128 --no-preprocessing 2>&1 | tail -1 | cut -c1-330
} > gpurun_out/r2q_stream_check.log 2>&1
cat gpurun_out/r2q_stream_check.log
