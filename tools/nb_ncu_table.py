"""Prints the ncu metric CSV of tools/gpu_nb_ncu.sh as one line per launch."""
import csv, sys, collections
for path in sys.argv[1:]:
    rows = [l for l in open(path) if not l.startswith("==")]
    d = collections.OrderedDict()
    for r in csv.DictReader(rows):
        d.setdefault((r["ID"], r["Kernel Name"][:60]), {})[r["Metric Name"]] = r["Metric Value"]
    print("#", path)
    for (i, k), m in d.items():
        t = float(m["gpu__time_duration.sum"].replace(",", "")) / 1e3
        print("%3s %-60s %7.2f us  inst %9s  issue %5s%%  warps %5s%%  regs %3s grid %s" % (
            i, k, t, m["smsp__inst_executed.sum"], m["smsp__issue_active.avg.pct_of_peak_sustained_active"][:5],
            m["sm__warps_active.avg.pct_of_peak_sustained_active"][:5], m["launch__registers_per_thread"], m["launch__grid_size"]))
