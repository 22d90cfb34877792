"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=45):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
        tot[name] += v
        cnt[name] += 1
        n += 1
    total = sum(tot.values())
    print("total %.1f us over %d launches (cold-cache, serialised: compare shares)" % (total, n))
    ours = sum(v for k, v in tot.items() if "effimvs" in k)
    print("libeffimvs kernels: %.1f us (%.1f%%)" % (ours, 100 * ours / total))
    for k, v in sorted(tot.items(), key=lambda x: -x[1])[:top]:
        print("%9.1f us %5.1f%% x%3d  %s" % (v, 100 * v / total, cnt[k], k))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
