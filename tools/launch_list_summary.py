#!/usr/bin/env python
"""Per-kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list (profiles/)."""
import collections
import csv
import sys


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(unit, 1e-6)
        name = r[ix["Kernel Name"]]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    print(f"# ncu launch list `{sys.argv[1]}`: {sum(cnt.values())} launches, {T:.2f} ms in total (cold-cache, serialised)\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"| `{k[:100]}` | {cnt[k]} | {v:.3f} | {100 * v / T:.1f} % |")


if __name__ == "__main__":
    main()
