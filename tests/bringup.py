"""GPU bring-up driver: runs every gpu test file in its own process (a device trap in
one cannot poison the next), writes logs to gpurun_out/.  Usage on the GPU box:
    python tests/bringup.py [pytest -k expression]
"""
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)
files = ["tests/test_gpu_seg.py", "tests/test_gpu_unet.py"]
rc_all = 0
for f in files:
    t0 = time.time()
    cmd = [sys.executable, "-m", "pytest", f, "-m", "gpu", "-q", "-s", "-x" if os.environ.get("SD_X") else "-q",
           "--timeout", "600", "-p", "no:cacheprovider"] + (["-k", " ".join(sys.argv[1:])] if len(sys.argv) > 1 else [])
    log = OUT / (Path(f).stem + ".log")
    with open(log, "w") as fh:
        try:
            rc = subprocess.run(cmd, cwd=ROOT, stdout=fh, stderr=subprocess.STDOUT, timeout=900).returncode
        except subprocess.TimeoutExpired:
            rc = 124
    rc_all |= rc
    tail = "".join(open(log).readlines()[-25:])
    print(f"==== {f}: rc={rc} in {time.time() - t0:.0f}s\n{tail}")
sys.exit(1 if rc_all else 0)
