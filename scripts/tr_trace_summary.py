"""Summarise an NB200_TR_TRACE file: mean ns between consecutive phase boundaries of CTA 0, by tag
(100+l: FWD l, 200: LOSS, 300+l: BWD l, 400: REDUCE, 500: ADAM, 600: epoch end incl. validation)."""
import sys, collections
d = collections.defaultdict(list); prev = None
for line in open(sys.argv[1]):
    if line.startswith("#"): prev = None; continue
    tag, t = map(int, line.split())
    if prev is not None: d[tag].append(t - prev)
    prev = t
tot = 0
for tag in sorted(d):
    v = d[tag][len(d[tag]) // 4:]  # skip the cold start
    m = sum(v) / len(v); print(f"{tag:4d} n={len(v):5d} mean {m/1e3:7.2f} us  min {min(v)/1e3:7.2f}")
    if tag < 600: tot += m
print(f"step (sum of phase means): {tot/1e3:.1f} us")
