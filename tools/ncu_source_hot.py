"""Summarise an `ncu --page source --csv` export: executed-instruction and stall-sample totals by opcode class, and
the hottest SASS ranges.  usage: ncu -i rep --page source --csv > x.csv; python tools/ncu_source_hot.py x.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
by_op = collections.Counter(); samp_op = collections.Counter()
for r in data:
    src = r[ix["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0]
    by_op[op] += int(r[ix["Instructions Executed"]]); samp_op[op] += int(r[ix["# Samples"]])
print(f"total warp instructions {tot_inst}, samples {tot_samp}")
for op, n in by_op.most_common(25):
    print(f"  {op:10s} inst {n:10d} ({100*n/tot_inst:5.1f}%)  samples {samp_op[op]:7d} ({100*samp_op[op]/max(tot_samp,1):5.1f}%)")
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
st = {n: sum(int(r[ix[n]]) for r in data) for n in stalls}
print("stall samples:", ", ".join(f"{k[6:]}={v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v))
# hottest instructions by samples
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for r in top:
    dom = max(stalls, key=lambda n: int(r[ix[n]]))
    print(f"  {r[ix['# Samples']]:>6s} samples  inst {r[ix['Instructions Executed']]:>8s}  {dom[6:]:12s} {r[ix['Source']].strip()[:90]}")
