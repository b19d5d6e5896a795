"""Build the source-line -> block table for tools/ncu_by_line.py from the current sources."""
import json, re, sys
out = []
for f in ("cvr_device.cuh", "cvr_kernels.cuh"):
    lines = open(f"cudavolumerenderer_b200/csrc/{f}").read().splitlines()
    marks = []
    for i, l in enumerate(lines, 1):
        m = re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:CVR_DEV|__global__)\s+[\w:<>\s\*&]*?\b(\w+)\s*\(", l)
        if m and m.group(1) not in ("if", "for", "while"):
            marks.append((i, m.group(1)))
        m2 = re.match(r"\s*k_volpt(\w*)\(const __grid_constant__", l)
        if m2:
            marks.append((i, "kernel:k_volpt" + m2.group(1)))
    for j, (ln, name) in enumerate(marks):
        hi = marks[j + 1][0] - 1 if j + 1 < len(marks) else len(lines)
        out.append({"name": name, "file": f, "lo": ln, "hi": hi})
json.dump(out, open(sys.argv[1], "w"), indent=0)
