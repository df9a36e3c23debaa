"""Read an `ncu --page source --csv` dump and print the instructions with the most stall samples.
usage: ncu -i rep --page source --csv > src.csv; python tools/ncu_top.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
isrc, isamp, iexec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[isamp] or 0) for r in body)
print("total samples", tot)
order = sorted(range(len(body)), key=lambda k: -int(body[k][isamp] or 0))[:n]
for k in sorted(order):
    r = body[k]
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{k:5d} {int(r[isamp]):7d} ({100*int(r[isamp])/max(tot,1):5.1f}%) exec={r[iexec]:>8s} {r[isrc].strip()[:70]:70s} {st}")
