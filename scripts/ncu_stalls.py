"""Which instructions carry each stall reason (source page of an .ncu-rep)."""
import csv, subprocess, sys
rep = sys.argv[1]; reason = sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]; h = rows[hi]
ci, src = h.index(reason), h.index("Source")
data = []
for idx, r in enumerate(rows[hi + 1:]):
    try: data.append((int(r[ci]), idx, r[src].strip()))
    except (ValueError, IndexError): pass
tot = sum(d[0] for d in data)
print(reason, "total", tot)
for n, idx, s in sorted(data, reverse=True)[:top]:
    print(f"  {n:6d} {100*n/max(tot,1):5.1f}% #{idx:5d} {s[:90]}")
