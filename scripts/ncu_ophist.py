"""Opcode histogram (weighted by executed warp-instructions) and exec-count regions of an .ncu-rep."""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]; h = rows[hi]
src, ex = h.index("Source"), h.index("Instructions Executed")
hist = collections.Counter(); tot = 0
regions = []  # (exec, n_instr, first_idx)
for idx, r in enumerate(rows[hi + 1:]):
    try: e = int(r[ex])
    except (ValueError, IndexError): continue
    s = r[src].strip()
    toks = s.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = ".".join(op.split(".")[:2])
    hist[op] += e; tot += e
    if regions and regions[-1][0] == e: regions[-1][1] += 1
    else: regions.append([e, 1, idx])
print("total warp-instr", tot)
for op, n in hist.most_common(40): print(f"  {op:24s} {n:12d} {100*n/tot:5.1f}%")
print("regions (exec count x #instr) with >=1% of total:")
for e, n, i in regions:
    if e * n >= 0.01 * tot: print(f"  idx {i:5d}: exec={e:8d} x {n:4d} instr = {100*e*n/tot:5.1f}%")
