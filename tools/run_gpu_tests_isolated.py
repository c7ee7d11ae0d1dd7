"""Run every `-m gpu` test in its own process with a timeout (a trapped kernel poisons the CUDA context of
its process; isolation keeps the other results).  Writes gpurun_out/isolated_tests.log and exits non-zero if any failed."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)
os.makedirs("gpurun_out", exist_ok=True)
args = sys.argv[1:] or ["tests"]
r = subprocess.run([sys.executable, "-m", "pytest", "--collect-only", "-q", "-m", "gpu", *args], capture_output=True, text=True)
ids = [l.strip() for l in r.stdout.splitlines() if "::" in l]
print(f"collected {len(ids)} gpu tests", flush=True)
fails = 0
with open("gpurun_out/isolated_tests.log", "w") as log:
    for t in ids:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "--no-header", "-p", "no:cacheprovider", t],
                               capture_output=True, text=True, timeout=int(os.environ.get("PG_TEST_TIMEOUT", "240")))
            ok, out = p.returncode == 0, p.stdout[-3000:] + p.stderr[-1500:]
        except subprocess.TimeoutExpired as e:
            ok, out = False, f"TIMEOUT\n{(e.stdout or b'')[-2000:]}"
        dt = time.time() - t0
        line = f"{'PASS' if ok else 'FAIL'} {dt:6.1f}s {t}"
        print(line, flush=True)
        log.write(line + "\n")
        if not ok:
            fails += 1
            log.write(out + "\n" + "-" * 100 + "\n")
            print(out[-1500:], flush=True)
print(f"{len(ids) - fails} passed, {fails} failed")
sys.exit(1 if fails else 0)
