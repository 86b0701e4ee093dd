#!/usr/bin/env python
"""Run the GPU tests file by file; any node that fails (or errors) is re-run alone in a fresh
process, because a trapped kernel poisons its CUDA context and would take the rest of the file
down with it.  Writes gpurun_out/stage.log.

    python tools/gpu_stage.py [pytest args / files...]
"""
import re
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
PY = [sys.executable, "-m", "pytest", "-q", "--no-header", "-p", "no:cacheprovider", "-m", "gpu", "-rfE"]


def run(args, timeout):
    t0 = time.time()
    try:
        r = subprocess.run(PY + args, capture_output=True, text=True, cwd=ROOT, timeout=timeout)
        return r.returncode, r.stdout + r.stderr, time.time() - t0
    except subprocess.TimeoutExpired as e:
        return 124, "TIMEOUT\n" + str(e.stdout)[-3000:], time.time() - t0


def main():
    files = sys.argv[1:] or sorted(str(p.relative_to(ROOT)) for p in (ROOT / "tests").glob("test_gpu_*.py"))
    log = open(OUT / "stage.log", "w")
    bad_total = []
    for f in files:
        rc, text, dt = run([f], 900)
        log.write(f"===== {f}: rc={rc} ({dt:.1f}s)\n{text[-6000:]}\n")
        log.flush()
        if rc == 0:
            continue
        nodes = sorted(set(re.findall(r"^(?:FAILED|ERROR) (\S+)", text, flags=re.M)))
        if rc == 124 or not nodes:
            coll = subprocess.run(PY + ["--collect-only", f], capture_output=True, text=True, cwd=ROOT)
            nodes = [l.strip() for l in coll.stdout.splitlines() if "::" in l]
        for node in nodes:
            rc2, text2, dt2 = run(["-x", node], 300)
            log.write(f"----- alone: {'PASS' if rc2 == 0 else 'FAIL'} {node} ({dt2:.1f}s)\n")
            if rc2 != 0:
                bad_total.append(node)
                log.write(text2[-5000:] + "\n")
            log.flush()
    log.write("\nFAILING NODES:\n" + "\n".join(bad_total) + "\n")
    log.close()
    print(f"{len(bad_total)} failing nodes")
    for n in bad_total:
        print("FAIL", n)
    sys.exit(1 if bad_total else 0)


if __name__ == "__main__":
    main()
