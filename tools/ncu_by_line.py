"""Join an `ncu --page source --csv` export (per SASS instruction) with `nvdisasm --print-line-info` of the same cubin
and print executed warp-instructions / stall samples per CUDA source line (inlined callee lines are attributed to
the innermost file:line).  usage: python tools/ncu_by_line.py sass.csv dis.txt kernel_substring [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
kern = sys.argv[3]
lines = open(sys.argv[2]).read().split("\n")
# locate the kernel's text section
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
cur = ("?", 0); seq = []
for l in lines[start + 1:]:
    if l.startswith("//-------") or l.startswith("\t.section"):
        if seq: break
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        seq.append(cur)
assert len(seq) == len(data), (len(seq), len(data))
inst = collections.Counter(); samp = collections.Counter()
for loc, r in zip(seq, data):
    inst[loc] += int(r[ix["Instructions Executed"]]); samp[loc] += int(r[ix["# Samples"]])
ti, ts = sum(inst.values()), sum(samp.values())
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = {}
for loc, n in inst.most_common(top):
    f, ln = loc
    if f not in src:
        try: src[f] = open(next(p for p in ("wav2vec-s_b200/csrc/" + f, "tools/" + f) if __import__("os").path.exists(p))).read().split("\n")
        except StopIteration: src[f] = []
    text = src[f][ln - 1].strip()[:80] if 0 < ln <= len(src[f]) else ""
    print(f"{f}:{ln:<4d} inst {100*n/ti:5.1f}%  samples {100*samp[loc]/max(ts,1):5.1f}%  | {text}")
