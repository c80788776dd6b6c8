"""Geometry sweep of the specialised Q1 kernel: rows per thread x TMA stages x register budget.  Each variant is a different
generated kernel (NVRTC), measured in its own process.  usage: python scripts/jit_sweep.py [orders]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = sys.argv[1] if len(sys.argv) > 1 else "15000000"
for rpt in ("2", "1", "4"):
    for stages in ("2", "3", "4"):
        for minb in ("2", "3", "4"):
            env = dict(os.environ, SAILGPU_TIMING="1", SAILGPU_JIT_MIN_ROWS="0", SAILGPU_RPT=rpt, SAILGPU_JIT_STAGES=stages, SAILGPU_JIT_MINB=minb)
            try:
                r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "jit_diag.py"), "--one", n], env=env, capture_output=True, text=True, timeout=120)
                line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip()[-300:]
                try:
                    d = json.loads(line)
                    print(f"rpt={rpt} stages={stages} minb={minb}: kernel ms {d['kernel_ms_1']} {d['kernel_ms_2']}  jit={d['jit_launches']}", flush=True)
                except Exception:
                    print(f"rpt={rpt} stages={stages} minb={minb}: {line}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"rpt={rpt} stages={stages} minb={minb}: TIMEOUT", flush=True)
