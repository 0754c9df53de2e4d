"""Development helper: per-kernel table (avg ms per launch, share) of one or more bench.py JSON lines, side by side."""
import json
import sys

runs = []
for path in sys.argv[1:]:
    line = [l for l in open(path) if l.startswith("{")][-1]
    runs.append((path, json.loads(line)))
names = []
for _, r in runs:
    for k in r["roofline"]["kernels"]:
        if k not in names:
            names.append(k)
print(f"{'kernel':40s}" + "".join(f"{p[-28:]:>30s}" for p, _ in runs))
for k in names:
    row = f"{k:40s}"
    for _, r in runs:
        e = r["roofline"]["kernels"].get(k)
        per_frame = e.get("launches_per_frame", e.get("launches", 0) / r["steps"]) if e else 0.0
        row += f"{(e['avg_ms'] * per_frame) if e else float('nan'):30.4f}"
    print(row)
print(f"{'ms_per_step (device-resident)':40s}" + "".join(f"{r['ms_per_step']:30.4f}" for _, r in runs))
print(f"{'e2e ms_per_step':40s}" + "".join(f"{r['e2e']['ms_per_step']:30.4f}" for _, r in runs))
print(f"{'pairs_per_frame':40s}" + "".join(f"{r['config']['pairs_per_frame']:30.2f}" for _, r in runs))
print(f"{'candidates_per_frame':40s}" + "".join(f"{r['config']['candidates_per_frame']:30.2f}" for _, r in runs))
