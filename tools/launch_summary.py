"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share.
    python tools/launch_summary.py gpurun_out/launches.csv [top_n]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = row["Metric Unit"]
        ns = v * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += ns
        tot += ns
    print("# per-launch gpu__time_duration.sum (ncu --clock-control none; cold-cache, serialised: compare SHARES)")
    print("# total %.3f ms over %d launches" % (tot / 1e6, sum(a[0] for a in agg.values())))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%6d launches %10.3f ms  avg %9.2f us  %5.1f%%  %s" % (n, t / 1e6, t / n / 1e3, 100 * t / tot, k[-100:]))


if __name__ == "__main__":
    main()
