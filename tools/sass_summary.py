"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (tcgen05.mma = UTC*MMA, tcgen05.ld =
LDTM, TMA loads = UTMALDG, TMA stores = UTMASTG, cp.async.bulk = UBLKCP, legacy mma.sync = HMMA), from
`cuobjdump -sass` of the built library.  Writes profiles/<tag>_sass_summary.txt.

    python tools/sass_summary.py [tag]
"""
import re
import subprocess
import sys
from collections import OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "infodiffusion_b200" / "libidf_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "MUFU", "total"]

tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
kernels = OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), dict.fromkeys(KEYS, 0))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["total"] += 1
        for k in KEYS[:-1]:
            if op.startswith(k):
                cur[k] += 1
out = [f"SASS mnemonic counts per kernel of {LIB.name} (cuobjdump -sass, sm_100a); tcgen05.mma = UTCHMMA, tcgen05.ld = LDTM,",
       "TMA load / store = UTMALDG / UTMASTG, cp.async.bulk = UBLKCP, mbarrier ops = SYNCS, legacy tensor path = HMMA (must be 0)", "",
       f"{'kernel':74s} " + " ".join(f"{k:>7s}" for k in KEYS)]
for name, c in kernels.items():
    d = re.sub(r"\((int|bool|unsigned int)\)", "", demangle(name))
    d = re.sub(r"\((idf::|const|float|int|unsigned|long|void|CUtensorMap|__nv).*$", "", d).replace("void idf::", "").replace("idf::", "")
    out.append(f"{d[:74]:74s} " + " ".join(f"{c[k]:7d}" for k in KEYS))
text = "\n".join(out) + "\n"
(ROOT / "profiles" / f"{tag}_sass_summary.txt").write_text(text)
print(text)
