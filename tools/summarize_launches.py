"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: launches, total and share of the
captured device time.  Usage: python tools/summarize_launches.py gpurun_out/launches.csv [steps_captured] > profiles/..."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in r:
        v = float(row[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(row[ui], 1e-6)
        name = row[ki].split("(")[0][:70]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v for _, v in agg.values())
    print("# %s: %d kernels, %.3f ms of device time captured (%.1f steps: %.3f ms/step; cold-cache serialised launches)" %
          (path, sum(c for c, _ in agg.values()), tot, steps, tot / steps))
    print("%-70s %8s %12s %12s %7s" % ("kernel", "launches", "total ms", "ms/step", "share"))
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %8d %12.3f %12.3f %6.1f%%" % (k, c, v, v / steps, 100 * v / tot))


if __name__ == "__main__":
    main()
