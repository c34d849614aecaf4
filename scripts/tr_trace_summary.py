"""Summarise an NB200_TR_TRACE file: mean ns between consecutive marks of CTA 0, keyed by
(previous tag -> tag).  Coarse tags: 100+l after FWD l, 200 LOSS, 300+l BWD l, 400 REDUCE, 500 ADAM,
600 epoch end; fine tags (a -DNB200_TR_FINE build): 1001.. inside FWD, 2001.. inside BWD, 4001.. REDUCE."""
import sys, collections
d = collections.OrderedDict(); prev = None
for line in open(sys.argv[1]):
    if line.startswith("#"): prev = None; continue
    tag, t = map(int, line.split())
    if prev is not None: d.setdefault((prev[0], tag), []).append(t - prev[1])
    prev = (tag, t)
coarse = collections.defaultdict(float)
for (a, b), v in d.items():
    v = v[len(v) // 4:]  # skip the cold start
    m = sum(v) / len(v)
    print(f"{a:5d} -> {b:5d}  n={len(v):5d}  mean {m/1e3:7.2f} us  min {min(v)/1e3:7.2f}")
