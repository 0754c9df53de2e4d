"""Opcode histogram of the hot kernels' SASS (cuobjdump -sass libc2d_gpu.so) -> profiles/<round>_sass_top_kernels.md.
Shows what the kernels are made of (no tensor-core / TMA opcodes are expected on this path: nothing is a dense
contraction; the traversal is LDG/ISETP/FSETP + shared-memory staging, the polygon kernels FMUL/FADD + SHFL/VOTE).

    python scripts/sass_summary.py colli2de_b200/libc2d_gpu.so profiles/r02_sass_top_kernels.md
"""
import collections
import re
import subprocess
import sys

KERNELS = ["k_traverse_sync", "k_narrow_manifold_pp", "k_narrow_filter_pp", "k_narrow_filter", "k_narrow_manifold", "k_refit_dirty",
           "k_fit", "k_apply_translates", "k_bin_scatter", "k_slab_mark", "k_slab_select", "k_sort_pass", "k_ray_block"]


def main(lib, out):
    text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    blocks = re.split(r"\n\s*Function : ", text)
    rows = []
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        short = re.sub(r"\(.*", "", dem).replace("c2d_dev::", "")
        if not any(k in short for k in KERNELS):
            continue
        ops = collections.Counter()
        for line in b.splitlines():
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m:
                ops[m.group(1).split(".")[0]] += 1
        total = sum(ops.values())
        top = ", ".join(f"{op} {n}" for op, n in ops.most_common(14))
        special = {k: ops[k] for k in ("SHFL", "VOTE", "MATCH", "ATOMS", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR",
                                        "MUFU", "HMMA", "UTMALDG", "UBLKCP", "LDGSTS") if ops[k]}
        rows.append((short, total, special, top))
    with open(out, "w") as f:
        f.write("# SASS opcode summary of the hot kernels (sm_100a, `cuobjdump -sass colli2de_b200/libc2d_gpu.so`)\n\n")
        f.write("No HMMA / UTMALDG / UBLKCP opcodes: the path has no dense contraction and no tile copies to stage — it is gathers,\n"
                "compares, warp votes / shuffles and atomics.  LDL / STL = local-memory traffic (traversal stacks, spills).\n\n")
        f.write("| kernel | instructions | memory / warp-level opcodes | most frequent opcodes |\n|---|---|---|---|\n")
        for short, total, special, top in sorted(rows, key=lambda r: -r[1]):
            f.write(f"| `{short}` | {total} | {', '.join(f'{k} {v}' for k, v in special.items())} | {top} |\n")
    print("wrote", out, len(rows), "kernels")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
