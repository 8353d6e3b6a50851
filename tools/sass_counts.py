"""SASS evidence: tensor-core / TMEM / TMA / cluster instructions per kernel of wav2vec-s_b200/lib/libw2vs.so (sm_100a).
    python tools/sass_counts.py > profiles/r02_sass_counts.txt
cuobjdump -sass on the library, mnemonics counted per `Function :`.  UTCHMMA = tcgen05.mma (.2CTA = cta_group::2),
LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, LDGSTS = cp.async,
HMMA = mma.sync, UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, UCGABAR = barrier.cluster, STAS = st.async
(remote shared-memory store with mbarrier completion), REDG = red.global."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wav2vec-s_b200", "lib", "libw2vs.so")
COLS = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "UTCBAR", "SYNCS",
        "UCGABAR", "STAS", "REDG"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
counts, total, cur = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for c in COLS:
            if op == c or op.startswith(c + ".") or (c == "UTCHMMA" and op.startswith("UTCHMMA") and ".2CTA" not in op):
                if c == "UTCHMMA" and ".2CTA" in op:
                    continue
                if c == "UTCHMMA.2CTA" and ".2CTA" not in op:
                    continue
                counts[cur][c] += 1
                break
print(__doc__.strip().replace("\n", "\n# ").join(["# ", ""]))
print("kernel".ljust(72) + "".join(c.rjust(len(c) + 2) for c in COLS) + "   instr")
for k in sorted(total):
    if not any(counts[k].values()) and "conv0" in k:
        continue        # the 40 template instances of the SIMT first-conv kernel carry none of these
    name = subprocess.run(["c++filt", "-p", k], capture_output=True, text=True).stdout.strip() or k
    name = re.sub(r"w2vs::\(anonymous namespace\)::|w2vs::", "", name)[:70]
    print(name.ljust(72) + "".join(str(counts[k][c]).rjust(len(c) + 2) for c in COLS) + str(total[k]).rjust(8))
