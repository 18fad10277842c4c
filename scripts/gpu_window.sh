#!/usr/bin/env bash
tag=${1:-window}
mkdir -p gpurun_out
{
  echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_cta.py tests/test_gpu_bnb.py "tests/test_refbuild.py::test_batched_cpp_node_loop_reaches_the_ip_optimum" -x -q 2>&1 | tail -6
  for inst in scpnre1 scpnrg1; do for extra in "--slots 128" "--slots 148" "--slots 296" "--slots 148 --warm-start"; do
    echo "== bnb $inst $extra"; timeout 600 python bench.py --workload bnb --bnb-instance $inst --steps 8 --warmup 3 $extra 2>>gpurun_out/${tag}.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['bnb']
print({a:(round(b[a],2) if isinstance(b[a],float) else b[a]) for a in ('value','nodes','lp_iterations_per_node','lp_device_ms_per_node','ms_per_round','incumbent')}); print('  ', b['rank0']['round_ms'])"
  done; done
  echo "== SB200_WINDOW_LAUNCH=0 (per-slot launches) scpnre1 128"; SB200_WINDOW_LAUNCH=0 timeout 600 python bench.py --workload bnb --bnb-instance scpnre1 --steps 8 --warmup 3 --slots 128 2>>gpurun_out/${tag}.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); b=d['bnb']; print(round(b['value'],1), b['rank0']['round_ms'])"
  python - <<'PY'
import sys, subprocess
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_golden
from oracle import scp_io
for nm in ("scpnre1", "scpnrg1"):
    inst, _ = load_golden(nm)
    scp_io.write_scp_text(inst, f"/tmp/{nm}.txt")
    for sl in ("128", "148"):
        r = subprocess.run(["oracle/_ref/bnb_batched_b200", f"/tmp/{nm}.txt", "--max-iter", "100", "--max-nodes", "3000", "--slots", sl, "--no-preprocessing"], capture_output=True, text=True, timeout=300)
        print("C++", nm, sl, r.stdout.strip()[-520:-260])
PY
  tail -3 gpurun_out/${tag}.err
} > gpurun_out/${tag}.log 2>&1
cat gpurun_out/${tag}.log
