"""Builds and runs the UMMA shifted-descriptor probe (tools/cu/shift_probe.cu, a standalone program: it is NOT part of
libidf_b200.so) for several row shifts and both base_offset conventions.  Needs a B200."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
exe = ROOT / "tools" / "cu" / "shift_probe"
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                str(ROOT / "tools" / "cu" / "shift_probe.cu"), "-o", str(exe), "-lcuda"], check=True)
sys.exit(subprocess.run([str(exe)]).returncode)
