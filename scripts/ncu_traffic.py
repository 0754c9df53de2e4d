"""Per-kernel averages of an ncu summary CSV (scripts/ncu_summary.py) -> the JSON bench.py reads for roofline.traffic.

    python scripts/ncu_traffic.py profiles/r01_ncu_full_top_kernels_v3.csv profiles/r01_ncu_traffic.json
"""
import csv
import json
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"^(c2d_dev::|c2d_host::)", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(?:<[^>]*>)?)", name)
    return m.group(1) if m else name


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr = [h.split(" [")[0] for h in rows[0]]
    units = [h.split(" [")[1].rstrip("]") if " [" in h else "" for h in rows[0]]
    acc = defaultdict(lambda: defaultdict(list))
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))

        def val(key, scale_units=True):
            v = float(d[key].replace(",", ""))
            unit = u.get(key, "")
            if scale_units:
                v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-3, "ns": 1e-6, "ms": 1.0, "ms ": 1.0}.get(unit, 1.0)
            return v
        k = acc[short(d["Kernel Name"])]
        k["dram"].append(val("dram__bytes_read.sum") + val("dram__bytes_write.sum"))
        k["ms"].append(val("gpu__time_duration.sum"))
        k["lanes"].append(val("smsp__thread_inst_executed_per_inst_executed.ratio", False))
        k["l1"].append(val("l1tex__t_sector_hit_rate.pct", False))
        k["l2"].append(val("lts__t_sector_hit_rate.pct", False))
        k["occ"].append(val("sm__warps_active.avg.pct_of_peak_sustained_active", False))
        k["dram_pct"].append(val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False))
        k["sm_pct"].append(val("sm__throughput.avg.pct_of_peak_sustained_elapsed", False))
        k["regs"].append(val("launch__registers_per_thread", False))
    out = {}
    for name, k in acc.items():
        out[name] = {key: sum(v) / len(v) for key, v in k.items()}
        out[name]["launches_profiled"] = len(k["ms"])
    json.dump(out, open(dst, "w"), indent=1)
    print("wrote", dst, sorted(out))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
