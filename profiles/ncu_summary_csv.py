"""Summarise ncu captures exported to CSV on the GPU box (profiles/run_ncu_r02.sh): key metrics, stall reasons,
opcode mix and hottest SASS lines.   python profiles/ncu_summary_csv.py gpurun_out/ncu/<name> [units_per_launch]"""
import collections, csv, sys

csv.field_size_limit(1 << 30)
base = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
NAMES = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
         'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
         'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
         'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
         'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
         'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
         'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
         'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
         'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
         'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
         'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active']
rows = [r for r in csv.reader(open(base + "_raw.csv", errors="replace")) if r]
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr, unit, val = rows[start], rows[start + 1], rows[start + 2]
print("==", base.split("/")[-1], val[hdr.index("Kernel Name")][:120])
for n in NAMES:
    if n in hdr:
        i = hdr.index(n)
        print(f"  {n:72s} {val[i]} {unit[i]}")
st = sorted(((float(val[i].replace(",", "") or 0), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
             if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h), reverse=True)[:7]
print("  stalls:", ", ".join(f"{h} {int(v)}" for v, h in st))
try:
    rows = [r for r in csv.reader(open(base + "_source.csv", errors="replace")) if r]
except FileNotFoundError:
    rows = []
hi = next((i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r), None)
if hi is not None:
    h2 = rows[hi]
    si, ei, sm = h2.index('Source'), h2.index('Instructions Executed'), h2.index('# Samples')
    ops, tot, hot = collections.Counter(), 0, []
    for r in rows[hi + 1:]:
        try:
            n, s = int(r[ei].replace(",", "")), int(r[sm].replace(",", ""))
        except Exception:
            continue
        parts = r[si].strip().split()
        if not parts:
            continue
        op = (parts[1] if parts[0].startswith('@') and len(parts) > 1 else parts[0]).split('.')[0]
        ops[op] += n
        tot += n
        hot.append((s, n, r[si].strip()[:90]))
    if tot:
        per = f" ({tot / units:.1f} per unit)" if units else ""
        print(f"  warp instructions {tot}{per}: " + " ".join(f"{o}:{100 * n / tot:.1f}%" for o, n in ops.most_common(16)))
        tots = sum(x[0] for x in hot) or 1
        for s, n, t in sorted(hot, reverse=True)[:8]:
            print(f"    {100 * s / tots:5.1f}% samples  exec={n:<12d} {t}")
