"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv` output.
usage: ncu_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2]
isamp = hdr.index('# Samples'); isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed')
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[isamp]) for r in data)
print(f"total samples {tot}, instructions {len(data)}")
order = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:n]
for i in order:
    r = data[i]
    st = sorted(((int(r[c]), h) for c, h in stall_cols if r[c] not in ('', '0')), reverse=True)[:3]
    print(f"{i:5d} {int(r[isamp]):7d} {100*int(r[isamp])/tot:5.1f}%  ex={r[iex]:>9s}  {r[isrc].strip()[:70]:70s} {st}")
