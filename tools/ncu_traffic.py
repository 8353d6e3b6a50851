"""DRAM traffic of every kernel launch of one encoder step, per kernel class, from one ncu metrics pass.

On the GPU box (one GPU, after `python bench.py --steps 1 --warmup 3` has exited 0 without ncu):

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
        --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-incremental
    python tools/ncu_traffic.py gpurun_out/traffic.csv large_64x20s > profiles/r02_traffic_large_64x20s.json

The last `launches_per_step` launches before the end of the log are those of the profiled (per-class timing) step;
classes follow wav2vec-s_b200/cabi.py:KERNEL_CLASS.  bench.py reads the JSON for `roofline.traffic`."""
import csv
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main(path, workload):
    import importlib.util
    spec = importlib.util.spec_from_file_location("cabi", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                       "wav2vec-s_b200", "cabi.py"))
    cabi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cabi)
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    iname, imetric, ival = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iid = hdr.index("ID")
    launches = {}
    for r in rows[1:]:
        d = launches.setdefault(int(r[iid]), {"name": r[iname]})
        d[r[imetric]] = float(r[ival].replace(",", ""))
    ids = sorted(launches)
    names = [launches[i]["name"] for i in ids]
    # one step = from a conv0_kernel launch to just before the next one; take the last complete step
    starts = [k for k, n in enumerate(names) if "conv0_kernel" in n]
    if len(starts) < 2:
        raise SystemExit("need at least two steps in the log")
    lo, hi = starts[-2], starts[-1]
    per_class, counts, time_ns = {}, {}, {}
    for k in range(lo, hi):
        d = launches[ids[k]]
        mm = re.search(r"(\w+_kernel)", d["name"])
        cls = cabi.KERNEL_CLASS.get(mm.group(1) if mm else "", "rows")
        per_class[cls] = per_class.get(cls, 0.0) + d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        counts[cls] = counts.get(cls, 0) + 1
        time_ns[cls] = time_ns.get(cls, 0.0) + d.get("gpu__time_duration.sum", 0.0)
    print(json.dumps({"workload": workload, "bytes_per_step": {k: int(v) for k, v in per_class.items()},
                      "launches_per_step": counts, "serialised_ms_per_step": {k: v / 1e6 for k, v in time_ns.items()},
                      "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control "
                                "none, every launch of one step (cold-cache, serialised launches)"}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "large_64x20s")
