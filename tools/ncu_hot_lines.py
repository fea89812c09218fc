"""Top source lines of a kernel by warp-stall samples, from an ncu report captured with --set full --import-source on.
usage: ncu_hot_lines.py REPORT.ncu-rep [N]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None; lines = []
for r in rows:
    if r and r[0] in ("#", "Line", "Address") or (r and "# Samples" in r):
        hdr = r; continue
    if hdr and len(r) == len(hdr): lines.append(r)
if not hdr:
    print(out[:2000]); sys.exit(0)
ix = {h: i for i, h in enumerate(hdr)}
sc = "# Samples" if "# Samples" in ix else [h for h in hdr if "Samples" in h][0]
tot = sum(int(r[ix[sc]] or 0) for r in lines)
lines.sort(key=lambda r: -int(r[ix[sc]] or 0))
print("total samples", tot, "columns:", [h for h in hdr][:8])
for r in lines[:top]:
    print("%6.2f%%  %s" % (100.0 * int(r[ix[sc]] or 0) / max(1, tot), " | ".join(r[i] for i in range(min(3, len(r))))[:170]))
