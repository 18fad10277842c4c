#!/usr/bin/env bash
python - <<'PY'
import sys, subprocess, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import load_golden
from oracle import scp_io
inst, _ = load_golden("scpnre1")
scp_io.write_scp_text(inst, "/tmp/scpnre1.txt")
for wf in ("2", "1"):
    env = dict(os.environ, SB200_TRACE_WINDOWS="1")
    r = subprocess.run(["oracle/_ref/bnb_batched_b200", "/tmp/scpnre1.txt", "--max-iter", "100", "--max-nodes", "4000", "--slots", "148",
                        "--windows-in-flight", wf, "--no-preprocessing"], capture_output=True, text=True, timeout=300, env=env)
    print("== windows in flight", wf)
    lines = [l for l in r.stderr.splitlines() if l.startswith("[window]")]
    print("\n".join(lines[:6] + ["..."] + lines[-24:]))
    print([l for l in r.stdout.splitlines() if l.startswith("{")][-1][:300])
PY
