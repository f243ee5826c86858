"""Top SASS lines by warp-sampling count for one launch of an ncu report (needs --import-source on / -lineinfo)."""
import csv, io, subprocess, sys
rep, k, top = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
for i, r in enumerate(rows):
    if "Source" in r and "# Samples" in r:
        h, start = r, i + 1
        break
si, smp, ei = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall_cols = [(j, c) for j, c in enumerate(h) if c.startswith("stall_") or "Stall" in c]
data = []
for idx, r in enumerate(rows[start:]):
    if len(r) <= smp:
        continue
    try:
        data.append((int(r[smp] or 0), idx, r[si].strip(), int(r[ei] or 0), r))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "columns:", [c for c in h if "tall" in c][:12])
for s, idx, text, ex, r in sorted(data, key=lambda d: -d[0])[:top]:
    reasons = sorted(((int(r[j] or 0), c) for j, c in stall_cols if (r[j] or "0").isdigit() and int(r[j] or 0) > 0), reverse=True)[:2]
    print("%6d %5.1f%%  #%-5d ex=%-8d %-70s %s" % (s, 100.0 * s / max(tot, 1), idx, ex, text[:70], reasons))
