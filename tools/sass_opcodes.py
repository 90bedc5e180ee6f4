"""Per-kernel counts of the SASS opcodes that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md): cuobjdump -sass of the
in-tree library, grouped by kernel.     python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tae_b200", "libtae_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCBAR.2CTA.MULTICAST", "HMMA.16816", "MUFU.EX2",
       "FFMA2", "LDGSTS", "SYNCS", "REDG", "RED.E.ADD"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
counts, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if not m:
        continue
    op = m.group(1)
    counts[cur]["_total"] += 1
    for o in OPS:
        if op == o or op.startswith(o + "."):
            counts[cur][o] += 1
fp = subprocess.run([sys.executable, "-c", "from tae_b200 import build; print(build.fingerprint())"], capture_output=True, text=True,
                    cwd=ROOT).stdout.strip()
print(f"# cuobjdump -sass {os.path.relpath(lib, ROOT)}  (source fingerprint {fp[:16]}); opcode counts per kernel, zero columns omitted")
for name, c in counts.items():
    hits = {o: c[o] for o in OPS if c[o]}
    full = demangle(name)
    short = full[:full.index(">(") + 1] if ">(" in full else re.sub(r"\(.*", "", full)
    print(f"{short}: {c['_total']} instructions; " + (", ".join(f"{o} {n}" for o, n in hits.items()) if hits else "-"))
