"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python tools/agg_launches.py file.csv [first_id last_id]"""
import collections
import csv
import re
import sys


def main():
    f = sys.argv[1]
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    rows = list(csv.reader(open(f, errors="ignore")))
    hi_ = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi_]
    ki, vi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi_ + 1:]:
        if len(r) <= vi or not r[ii].isdigit() or not (lo <= int(r[ii]) < hi):
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"^void ", "", name)[:72]
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{f}: launches {sum(v[0] for v in agg.values())}, total {tot / 1e6:.3f} ms")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"  {v[1] / tot * 100:5.1f}%  n={v[0]:5d}  sum={v[1] / 1e6:8.3f} ms  avg={v[1] / v[0] / 1e3:9.1f} us  {k}")


if __name__ == "__main__":
    main()
