"""Per-source-line stall samples of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [kernel-substring] [top]"""
import csv, subprocess, sys
csv.field_size_limit(10**9)
rep = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ""; top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur, fn = [], None, ""
for r in rows:
    if r and r[0] in ("Function Name", "Kernel Name"): fn = r[1][:70]
    if r and r[0] == "Line No":
        cur = {"fn": fn, "hdr": r, "rows": []}; blocks.append(cur); continue
    if cur is not None and r and r[0].isdigit(): cur["rows"].append(r)
def I(x):
    try: return int(x)
    except Exception: return 0
for b in blocks:
    if sub not in b["fn"] or "# Samples" not in b["hdr"]: continue
    h = b["hdr"]; si = h.index("# Samples"); ie = h.index("Instructions Executed")
    tot = sum(I(r[si]) for r in b["rows"])
    if tot == 0: continue
    print(b["fn"], "lines", len(b["rows"]), "samples", tot)
    for r in sorted(b["rows"], key=lambda r: -I(r[si]))[:top]:
        print(f"{I(r[si])*100/tot:5.1f}% L{r[0]:>4} inst={r[ie]:>10} {r[1][:120]}")
