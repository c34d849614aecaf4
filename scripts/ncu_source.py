"""Top stalled SASS instructions of an .ncu-rep (source page)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]
si, src, ex = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
data = []
for idx, r in enumerate(rows[hi + 1:]):
    try:
        data.append((int(r[si]), idx, r[src].strip(), int(r[ex])))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for n, idx, s, e in sorted(data, key=lambda x: -x[0])[:top]:
    print(f"{n:6d} {100*n/tot:5.1f}%  #{idx:5d} exec={e:8d}  {s[:100]}")
