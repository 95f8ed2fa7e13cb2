#!/usr/bin/env python3
"""Per-phase (between BAR.SYNC) breakdown of an .ncu-rep source page.  Usage: ncu_phases.py rep nbatch"""
import csv, io, subprocess, sys
rep, nb = sys.argv[1], float(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[k], rows[k + 1:]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, key):
    try: return float(r[ix[key]])
    except Exception: return 0.0
seg, cur = [], []
for r in data:
    cur.append(r)
    if 'BAR.SYNC' in r[ix['Source']]: seg.append(cur); cur = []
seg.append(cur)
tot = sum(g(r, '# Samples') for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for i, s in enumerate(seg):
    smp = sum(g(r, '# Samples') for r in s); ins = sum(g(r, 'Instructions Executed') for r in s); wf = sum(g(r, 'L1 Wavefronts Shared') for r in s)
    dm = sum(g(r, 'Instructions Executed') for r in s if 'DMMA' in r[ix['Source']])
    fp = sum(g(r, 'Instructions Executed') for r in s if any(x in r[ix['Source']] for x in ('DFMA', 'DMUL', 'DADD')))
    mix = sorted(((sum(g(r, st) for r in s), st[6:]) for st in stalls), reverse=True)[:4]
    print(f"seg{i}: {s[0][ix['Address']][-5:]}-{s[-1][ix['Address']][-5:]} n={len(s):4d} samples={smp:6.0f} ({100 * smp / tot:4.1f}%) inst/batch={ins / nb:7.1f} "
          f"wf/batch={wf / nb:7.1f} dmma/batch={dm / nb:5.1f} dfp/batch={fp / nb:6.1f} | " + ", ".join(f"{n} {100 * v / max(smp, 1):.0f}%" for v, n in mix))
