#!/usr/bin/env python3
"""Digest one kernel launch of an .ncu-rep: key counters, stall mix, hottest SASS lines.  Usage: ncu_digest.py rep [ntop [launch]]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sel = ["--launch-skip", sys.argv[3], "--launch-count", "1"] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
        'sm__cycles_elapsed.avg.per_second', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor', 'smsp__inst_executed_pipe_fp64', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fp64_op_dmma']
for i, h in enumerate(hdr):
    if any(h == w or (w.endswith('tensor') and h.startswith(w)) or (w.endswith('dmma') and h.startswith(w)) for w in want):
        print(f"{h:88s} {vals[i]:>20s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[k], rows[k + 1:]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, key):
    try: return float(r[ix[key]])
    except Exception: return 0.0
tot = sum(g(r, '# Samples') for r in data)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print(f"\nsamples {tot:.0f}; stall mix:", ", ".join(f"{s[6:]} {100 * v / tot:.1f}%" for v, s in sorted(((sum(g(r, s) for r in data), s) for s in stalls), reverse=True)[:8]))
wf = sum(g(r, 'L1 Wavefronts Shared') for r in data); ex = sum(g(r, 'L1 Wavefronts Shared Excessive') for r in data)
print(f"shared wavefronts {wf:.0f}, excessive {ex:.0f}; instructions {sum(g(r, 'Instructions Executed') for r in data):.0f}")
print("\nhottest by samples:")
for r in sorted(data, key=lambda r: -g(r, '# Samples'))[:ntop]:
    st = sorted(((g(r, s), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"  {r[ix['Address']][-5:]} {r[ix['Source']].strip()[:58]:58s} smp={g(r, '# Samples'):6.0f} {st[0][1]}={st[0][0]:.0f} {st[1][1]}={st[1][0]:.0f}")
print("\nhottest by shared wavefronts:")
for r in sorted(data, key=lambda r: -g(r, 'L1 Wavefronts Shared'))[:ntop]:
    print(f"  {r[ix['Address']][-5:]} {r[ix['Source']].strip()[:58]:58s} exec={g(r, 'Instructions Executed'):8.0f} wf={g(r, 'L1 Wavefronts Shared'):9.0f} ideal={g(r, 'L1 Wavefronts Shared Ideal'):9.0f}")
