"""Per-kernel SASS instruction histogram of the shipped library (runs on any machine with the CUDA toolkit; no GPU):
the Blackwell-specific mnemonics that prove the code path -- UTCHMMA (tcgen05.mma), UTMALDG (TMA loads), LDTM/STTM (tensor
memory), FMNMX3 (3-input max) -- plus size and local-memory (spill) traffic of every instantiation.

    python scripts/sass_histogram.py > profiles/sass_histogram_r2.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "multi-modal_colpali_b200" / "_lib" / "liblis.so"
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.2CTA", "LDTM", "STTM", "UTCBAR", "SYNCS", "FMNMX3", "FMNMX", "LDL", "STL",
        "BRX", "CALL"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    cur, rows = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            rows[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            rows[cur]["_n"] += 1
            base = op.split(".")[0]
            rows[cur][base] += 1
            if base in ("UTCHMMA", "UTMALDG") and ".2CTA" in op:
                rows[cur][base + ".2CTA"] += 1
    print(f"# SASS histogram of {LIB.relative_to(ROOT)} (cuobjdump -sass); columns = instruction counts in the kernel image")
    print("# " + " ".join(f"{k:>12}" for k in ["instrs"] + KEYS) + "  kernel")
    tot = collections.Counter()
    for name, c in rows.items():
        tot.update(c)
        short = re.sub(r"CUtensorMap_st", "TMap", demangle(name))
        short = re.sub(r"\(.*", "", short)
        print("  " + " ".join(f"{c[k]:>12}" for k in ["_n"] + KEYS) + "  " + short)
    print("  " + " ".join(f"{tot[k]:>12}" for k in ["_n"] + KEYS) + "  TOTAL")


if __name__ == "__main__":
    sys.exit(main())
