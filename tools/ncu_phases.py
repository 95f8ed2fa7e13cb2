#!/usr/bin/env python3
"""Per-phase breakdown of an .ncu-rep source page of a stage kernel.
  ncu_phases.py rep                 warp-per-group / warp-pair kernel: phases cut at DMMA counts (volume | flux | epilogue)
  ncu_phases.py rep --bar nbatch    block-synchronous kernels (mma, ws): phases cut at BAR.SYNC, per-batch instruction counts"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[k], rows[k + 1:]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
def g(r, key):
    try: return float(r[ix[key]])
    except Exception: return 0.0
tot = sum(g(r, '# Samples') for r in data)
if len(sys.argv) > 2 and sys.argv[2] == "--bar":
    nb = float(sys.argv[3])
    seg, cur = [], []
    for r in data:
        cur.append(r)
        if 'BAR.SYNC' in r[ix['Source']]: seg.append(cur); cur = []
    seg.append(cur)
    for i, s in enumerate(seg):
        smp = sum(g(r, '# Samples') for r in s); ins = sum(g(r, 'Instructions Executed') for r in s)
        dm = sum(g(r, 'Instructions Executed') for r in s if 'DMMA' in r[ix['Source']])
        mix = sorted(((sum(g(r, st) for r in s), st[6:]) for st in stalls), reverse=True)[:4]
        print(f"seg{i}: n={len(s):4d} samples {100 * smp / tot:4.1f}% inst/batch={ins / nb:7.1f} dmma/batch={dm / nb:5.1f} | "
              + ", ".join(f"{n} {100 * v / max(smp, 1):.0f}%" for v, n in mix))
    sys.exit(0)
ndmma = sum(1 for r in data if 'DMMA' in r[ix['Source']])
# the volume contraction issues the first (KSV * VT) DMMAs: everything up to the first LIFT fragment; found as the largest
# DMMA-free gap after the first DMMA (descriptor decode and flux set-up sit between the two contractions)
pos = [i for i, r in enumerate(data) if 'DMMA' in r[ix['Source']]]
gaps = sorted(((pos[i + 1] - pos[i], i) for i in range(len(pos) - 1)), reverse=True)
nvol = gaps[0][1] + 1
cnt, ins, agg = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
n = 0
for r in data:
    if 'DMMA' in r[ix['Source']]: n += 1
    ph = '0 prologue + face set-up' if n == 0 else '1 volume' if n <= nvol else '2 flux + LIFT' if n < ndmma else '3 epilogue + tail'
    cnt[ph] += g(r, '# Samples'); ins[ph] += g(r, 'Instructions Executed')
    for st in stalls: agg[ph][st[6:]] += g(r, st)
print(f"{ndmma} DMMA sites, {nvol} in the volume contraction")
for ph in sorted(cnt):
    top = ", ".join(f"{k} {100 * v / max(cnt[ph], 1):.0f}%" for k, v in agg[ph].most_common(6))
    print(f"{ph:26s} samples {100 * cnt[ph] / tot:5.1f}%  instr {ins[ph] / 1e6:6.2f} M | {top}")
