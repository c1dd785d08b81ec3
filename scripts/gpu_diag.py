"""Run one pytest selection and summarise the kernel's mbarrier-timeout diagnostics by (block, warp, barrier, parity)."""
import collections, re, subprocess, sys
sel = sys.argv[1]
r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-q", "-x", "-p", "no:cacheprovider", "-k", sel, "-s"],
                   capture_output=True, text=True, errors="replace")
c = collections.Counter()
other = []
for l in (r.stdout + r.stderr).splitlines():
    if "lis:" not in l:
        continue
    m = re.search(r"block (\d+) thread (\d+) bar smem (0x[0-9a-f]+) parity (\d+)", l)
    if m:
        c[(int(m.group(1)), int(m.group(2)) // 32, m.group(3), m.group(4))] += 1
    else:
        other.append(l.strip()[:200])
for l in other[:10]:
    print(l)
for k in sorted(c)[:80]:
    print(k, c[k])
print("exit", r.returncode, (r.stdout + r.stderr)[-300:])
