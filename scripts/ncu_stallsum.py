"""Totals of each warp-stall reason for one kernel of an .ncu-rep (source page)."""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "-k", f"regex:{kern}", "--launch-skip", skip, "--launch-count", "1", "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]; h = rows[hi]
cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
tot = {h[i]: 0 for i in cols}
for r in rows[hi + 1:]:
    if len(r) <= max(cols): continue
    for i in cols:
        try: tot[h[i]] += int(r[i])
        except ValueError: pass
s = sum(tot.values())
print(kern, "samples", s)
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    if v: print(f"  {k:28s} {v:7d} {100*v/max(s,1):5.1f}%")
