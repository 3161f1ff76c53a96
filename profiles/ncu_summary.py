"""Summarise an .ncu-rep (key metrics, stall reasons, opcode mix, hottest SASS lines) as text.
    python profiles/ncu_summary.py report.ncu-rep [units_per_launch]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
NAMES = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
         'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
         'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
         'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
         'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
         'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
         'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
         'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
         'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active']
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit, val = rows[0], rows[1], rows[2]
print("==", rep, val[hdr.index("Kernel Name")][:110] if "Kernel Name" in hdr else "")
for n in NAMES:
    if n in hdr:
        i = hdr.index(n)
        print(f"  {n:72s} {val[i]} {unit[i]}")
st = sorted(((float(val[i] or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
             if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h), reverse=True)[:7]
print("  stalls:", ", ".join(f"{h} {int(v)}" for v, h in st))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
if len(rows) > 2:
    h2 = rows[1]
    si, ei, sm = h2.index('Source'), h2.index('Instructions Executed'), h2.index('# Samples')
    ops, tot, hot = collections.Counter(), 0, []
    for r in rows[2:]:
        try:
            n, s = int(r[ei]), int(r[sm])
        except Exception:
            continue
        parts = r[si].strip().split()
        op = (parts[1] if parts[0].startswith('@') else parts[0]).split('.')[0]
        ops[op] += n
        tot += n
        hot.append((s, n, r[si].strip()[:80]))
    per = f" ({tot / units:.1f} per unit)" if units else ""
    print(f"  warp instructions {tot}{per}: " + " ".join(f"{o}:{100 * n / tot:.1f}%" for o, n in ops.most_common(16)))
    tots = sum(x[0] for x in hot) or 1
    for s, n, t in sorted(hot, reverse=True)[:8]:
        print(f"    {100 * s / tots:5.1f}% samples  exec={n:<12d} {t}")
