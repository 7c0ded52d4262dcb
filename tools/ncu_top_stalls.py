"""Top stalled SASS instructions of one launch in an .ncu-rep (source page).  usage: ncu_top_stalls.py rep skip [n]"""
import csv
import subprocess
import sys

rep, skip = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 20
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
print(rows[0][1][:110])
hdr = rows[1]
iS, iSrc = hdr.index("# Samples"), hdr.index("Source")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) > iS and r[iS].isdigit()]
tot = sum(int(r[iS]) for r in body)
print("total samples", tot)
agg = {}
for r in body:
    for i in stalls:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
body.sort(key=lambda r: -int(r[iS]))
for r in body[:n]:
    st = sorted([(int(r[i] or 0), hdr[i]) for i in stalls], reverse=True)[:2]
    print(r[iS].rjust(6), f"{100 * int(r[iS]) / tot:5.1f}%", r[iSrc].strip()[:64].ljust(64), st)
