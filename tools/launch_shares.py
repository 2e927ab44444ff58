#!/usr/bin/env python
"""Per-kernel totals and shares out of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file list.csv ...`).
    python tools/launch_shares.py list.csv ["title line"]"""
import collections
import csv
import re
import sys


def main():
    rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        name = re.sub(r"^(void )?(msfm::)?", "", r[ik]).split("(")[0]
        ns = float(r[iv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[iu], 1.0)
        tot[name] += ns
        cnt[name] += 1
    total = sum(tot.values())
    if len(sys.argv) > 2:
        print(sys.argv[2])
    for name, ns in tot.most_common():
        print(f"{name:<64s} {cnt[name]:4d} launches {ns / 1e6:10.3f} ms  share {100 * ns / total:5.1f} %")


if __name__ == "__main__":
    main()
