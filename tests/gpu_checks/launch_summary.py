"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list (one bench step).
Usage: python tests/gpu_checks/launch_summary.py launches.csv out.csv ["comment"]"""
import csv
import re
import sys
from collections import defaultdict


def main(src, out, comment=""):
    rows = [r for r in csv.reader(open(src)) if r]
    while rows and "Kernel Name" not in rows[0]:
        rows.pop(0)
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    acc = defaultdict(lambda: [0, 0.0])
    n = 0
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = re.sub(r"^void ", "", r[ik].split("(")[0])
        v = float(r[iv].replace(",", ""))
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(r[iu], 1e-6)
        acc[name][0] += 1
        acc[name][1] += ms
        n += 1
    tot = sum(v[1] for v in acc.values())
    with open(out, "w", newline="") as f:
        if comment:
            f.write("# " + comment + "\n")
        f.write(f"# {n} kernel launches, serialized cold-cache sum {tot:.3f} ms\n")
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "ms", "share"])
        for k, (c, ms) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, c, f"{ms:.4f}", f"{ms / tot:.4f}"])
    print(n, "launches", f"{tot:.3f} ms")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
