"""Top stall-sample SASS lines per kernel from `ncu -i X --page source --csv` output.
usage: ncu_hot.py file.csv [section_index] [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
want = int(sys.argv[2]) if len(sys.argv) > 2 else None
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
for k, s in enumerate(starts):
    e = starts[k + 1] if k + 1 < len(starts) else len(rows)
    hdr = rows[s + 1]
    body = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
    iS = hdr.index('Source'); iN = hdr.index('# Samples'); iE = hdr.index('Instructions Executed')
    stall = [i for i, x in enumerate(hdr) if x.startswith('stall_') and 'Not Issued' not in x]
    tot = sum(int(r[iN] or 0) for r in body)
    print('== section %d: %s | samples %d, lines %d' % (k, rows[s][1][:70], tot, len(body)))
    if want is None or want != k:
        continue
    top = sorted(body, key=lambda r: -int(r[iN] or 0))[:n]
    for r in top:
        st = sorted(((int(r[i] or 0), hdr[i]) for i in stall), reverse=True)[:2]
        print('%6d %5.1f%% exec=%-8s %-70s %s' % (int(r[iN]), 100.0 * int(r[iN]) / max(tot, 1), r[iE], r[iS].strip()[:70], st))
