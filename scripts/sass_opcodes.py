"""Per-kernel SASS opcode histogram of the shipped library: the mnemonics that prove which hardware
paths a kernel uses (B200_PROFILING.md: UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st,
UTCBAR = tcgen05.commit, UTMALDG / UBLKCP = TMA, LDGSTS = cp.async, HMMA = legacy mma.sync, FFMA2 = packed fp32).
    python scripts/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nessai_b200", "lib", "libnessai_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "HMMA", "FFMA2", "FFMA", "DFMA", "MUFU", "F2FP", "BAR", "SYNCS", "ATOM", "RED", "LDS", "STS", "LDG", "STG"]
cur, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1).split(".")[0]
        hist[cur][op] += 1
        hist[cur]["_total"] += 1
print(f"# SASS opcode counts per kernel of {os.path.basename(lib)} (cuobjdump -sass; static instruction counts)")
print("kernel," + ",".join(WATCH) + ",total")
for k, h in hist.items():
    name = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"\"{name}\"," + ",".join(str(h.get(w, 0)) for w in WATCH) + f",{h['_total']}")
