#!/usr/bin/env bash
# the C++ node loop on the C ABI (integration/sypha_bnb_batched_b200.cpp, linked against the reference's objects):
# its tests, then nodes/s with one window at a time and with two in flight
mkdir -p gpurun_out
{
  echo "== pytest refbuild"; timeout 600 python -m pytest tests/test_refbuild.py -m gpu -x -q 2>&1 | tail -4
  python - <<'PY'
import sys, subprocess, json
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_golden
from oracle import scp_io
for nm in ("scpnre1", "scpnrg1"):
    inst, _ = load_golden(nm)
    scp_io.write_scp_text(inst, f"/tmp/{nm}.txt")
    for wf in ("1", "2", "2", "1"):
        r = subprocess.run(["oracle/_ref/bnb_batched_b200", f"/tmp/{nm}.txt", "--max-iter", "100", "--max-nodes", "8000", "--slots", "148",
                            "--windows-in-flight", wf, "--no-preprocessing"], capture_output=True, text=True, timeout=300)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        try:
            d = json.loads(line[-1])
            print("C++", nm, "windows in flight", wf, {k: d.get(k) for k in ("objective", "nodes", "lp_iterations", "nodes_per_sec", "wall_ms", "base_cols")})
        except Exception as e:
            print("C++", nm, wf, "no result", repr(e), r.stdout[-300:], r.stderr[-300:])
PY
} > gpurun_out/cpp_loop.log 2>&1
cat gpurun_out/cpp_loop.log
