"""Per-kernel counts of the SASS mnemonics that show which hardware paths the shipped library uses (tcgen05: UTCHMMA / LDTM /
UTCBAR, TMA: UTMALDG, mbarrier: SYNCS, FP64 tensor cores: DMMA, FP64 pipe: DFMA / DADD / DMUL, programmatic dependent
launch: ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents).  Runs without a GPU.
usage: python tools/sass_counts.py [library.so] > profiles/r2_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ch-bin_b200", "libchbin_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "ACQBULK", "PREEXIT", "LDG", "STG", "LDS", "STS", "ATOMG", "SHFL"]
op = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)")
cnt, cur = collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    m = op.match(line)
    if m and cur:
        cnt[cur][m.group(1)] += 1
        cnt[cur]["total"] += 1
names = list(cnt)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
def short(s):
    s = s.replace("(anonymous namespace)::", "")
    s = re.sub(r"^void ", "", s)
    depth, out = 0, ""
    for ch in s:  # cut the parameter list, keep the template arguments
        if ch == "<": depth += 1
        if ch == ">": depth -= 1
        if ch == "(" and depth == 0: break
        out += ch
    return out
print("%-36s" % "kernel" + "".join("%8s" % k for k in KEYS) + "   total")
tot = collections.Counter()
for f, d in sorted(zip(names, dem), key=lambda t: -cnt[t[0]]["total"]):
    c = cnt[f]
    print("%-36s" % short(d)[:36] + "".join("%8d" % c.get(k, 0) for k in KEYS) + "%8d" % c["total"])
    for k in KEYS + ["total"]:
        tot[k] += c.get(k, 0)
print("%-36s" % f"ALL ({len(names)} kernels)" + "".join("%8d" % tot[k] for k in KEYS) + "%8d" % tot["total"])
