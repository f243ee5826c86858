"""For every igemm launch in an ncu report: duration, tensor %, and how often each role spun on each barrier
(executed counts of the mbarrier try_wait loops, by shared-memory offset), which names the bottleneck."""
import csv, subprocess, sys, re, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
names = {0x30000: "full(MMA waits TMA)", 0x30080: "empty(TMA waits MMA)", 0x30100: "tmem_full(epi waits MMA)", 0x30120: "tmem_empty(MMA waits epi)"}
for k, r in enumerate(rows[2:]):
    dur = r[col["gpu__time_duration.sum"]]
    tp = r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]]
    l2 = r[col["lts__throughput.avg.pct_of_peak_sustained_elapsed"]]
    dr = r[col["dram__bytes_read.sum"]]; dw = r[col["dram__bytes_write.sum"]]
    print("launch %d: %s us  tensor %s%%  L2 %s%%  dram rd %s wr %s MB" % (k, dur, tp[:5], l2[:5], dr[:7], dw[:7]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src)))
    h = None
    for i, sr in enumerate(srows):
        if "Source" in sr and "Instructions Executed" in sr:
            h = sr; start = i + 1; break
    if h is None:
        continue
    si, ei, smp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    tot = 0
    waits = {}
    for sr in srows[start:]:
        if len(sr) <= ei: continue
        tot += int(sr[ei] or 0)
        m = re.search(r"SYNCS\.PHASECHK\.TRANS64\.TRYWAIT\s+\w+, \[(\w+)\+URZ\+0x([0-9a-f]+)\]", sr[si])
        if m:
            off = int(m.group(2), 16)
            key = None
            for base, nm in names.items():
                if base <= off < base + (0x80 if base < 0x30100 else 0x10):
                    key = nm
            waits[key or hex(off)] = waits.get(key or hex(off), 0) + int(sr[ei] or 0)
        m2 = re.search(r"SYNCS\.PHASECHK\.TRANS64\.TRYWAIT\s+\w+, \[(\w+)(\+0x([0-9a-f]+))?\]", sr[si])
        if m2 and not m:
            waits["dyn:" + sr[si].strip()[-28:]] = waits.get("dyn:" + sr[si].strip()[-28:], 0) + int(sr[ei] or 0)
    print("   warp-instructions executed %d; try_wait executions: %s" % (tot, {k2: v for k2, v in sorted(waits.items(), key=lambda kv: -kv[1])}))
