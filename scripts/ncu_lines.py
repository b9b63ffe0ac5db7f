"""Per-source-line stall samples of one kernel in an ncu report (needs -lineinfo):
    python scripts/ncu_lines.py report.ncu-rep [kernel-index] [min-percent]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
secs = []
cur = None
fpath = ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1]
    elif r[0] == "Function Name":
        cur = {"name": r[1], "file": fpath, "rows": [], "hdr": None}
        secs.append(cur)
    elif r[0] == "Line No" and cur is not None:
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(r)
for i, k in enumerate(secs):
    print(i, k["file"].split("/")[-1], k["name"][:60], len(k["rows"]))
k = secs[kidx]
h = k["hdr"]
si = h.index("# Samples")
ie = h.index("Instructions Executed")
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
lines = [r for r in k["rows"] if r[2] == "-"]          # source-line aggregate rows
tot = sum(int(r[si] or 0) for r in lines)
print(k["name"][:100], "samples", tot, "warp instr", sum(int(r[ie] or 0) for r in lines))
agg = {}
for r in lines:
    for i in stall:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
print(sorted(((v, k2) for k2, v in agg.items()), reverse=True)[:8])
for r in lines:
    s = int(r[si] or 0)
    if s >= tot * minpct / 100:
        st = sorted(((int(r[i] or 0), h[i][6:]) for i in stall), reverse=True)[:3]
        print("%5s %-100s | %5d %5.1f%% instr %8s %s" % (r[0], r[1].strip()[:100], s, 100 * s / tot, r[ie], st))
