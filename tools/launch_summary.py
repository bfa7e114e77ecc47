"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: total ms, launches, share."""
import collections
import csv
import sys


def summarize(path):
    hdr, agg = None, collections.OrderedDict()
    for r in csv.reader(open(path, errors="ignore")):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v, u = float(d["Metric Value"].replace(",", "")), d["Metric Unit"]
        ms = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[u] * v
        a = agg.setdefault(d["Kernel Name"].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += ms
    return agg


if __name__ == "__main__":
    agg = summarize(sys.argv[1])
    tot = sum(t for _, t in agg.values())
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t:10.3f} ms {100 * t / tot:5.1f}% {c:6d} launches  {t / c * 1e3:9.1f} us/launch  {n}")
    print(f"{tot:10.3f} ms total")
