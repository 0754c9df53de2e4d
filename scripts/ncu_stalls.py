"""Scheduler / warp-state view of an .ncu-rep (ncu --set full): per kernel, averaged over its profiled launches — issue slot
utilisation, eligible / active warps per scheduler and the warp-stall reasons (cycles a warp waits per issued instruction).

    python scripts/ncu_stalls.py gpurun_out/r02_prof_top.ncu-rep profiles/r02_ncu_stalls.csv
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict

KEEP = ["smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"^(c2d_dev::|c2d_host::)", "", name)
    m = re.match(r"([A-Za-z0-9_:]+(?:<[^>]*>)?)", name)
    return m.group(1) if m else name


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    cols = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if STALL.match(h)]
    acc = defaultdict(lambda: defaultdict(list))
    for r in rows[2:]:
        k = acc[short(r[cols["Kernel Name"]])]
        for h in KEEP + stalls:
            if h in cols and r[cols[h]] not in ("", "n/a"):
                k[h].append(float(r[cols[h]].replace(",", "")))
    names = [STALL.match(h).group(1) for h in stalls]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "issue slots busy per cycle", "eligible warps / scheduler", "active warps / scheduler",
                    "active lanes per instruction", "SM throughput %", "DRAM throughput %", "warps active % of peak",
                    "top stall reasons (warp-cycles stalled per issued instruction)"] + [f"stall_{n}" for n in names])
        for kernel, k in sorted(acc.items()):
            mean = lambda h: sum(k[h]) / len(k[h]) if k[h] else float("nan")
            st = {n: mean(h) for n, h in zip(names, stalls)}
            top = "; ".join(f"{n} {v:.2f}" for n, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
            w.writerow([kernel, len(k[KEEP[0]])] + [f"{mean(h):.3f}" for h in KEEP] + [top] + [f"{st[n]:.3f}" for n in names])
    print("wrote", out, len(acc), "kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
