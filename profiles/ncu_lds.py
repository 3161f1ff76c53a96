"""Per-instruction shared-memory wavefronts of an .ncu-rep: python profiles/ncu_lds.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = rows[1]
print("columns:", [c for c in h if "avefront" in c or "Shared" in c or "Conflict" in c])
si, ei = h.index("Source"), h.index("Instructions Executed")
wcol = [i for i, c in enumerate(h) if c.strip() == "L1 Wavefronts Shared"]
icol = [i for i, c in enumerate(h) if c.strip() == "L1 Wavefronts Shared Ideal"]
if not wcol:
    sys.exit("no shared wavefront column")
wi, ii = wcol[0], (icol[0] if icol else None)
items, tot = [], 0
for r in rows[2:]:
    try:
        w = int(r[wi]); n = int(r[ei])
    except Exception:
        continue
    tot += w
    if w:
        items.append((w, n, int(r[ii]) if ii is not None and r[ii] else -1, r[si].strip()))
items.sort(reverse=True)
print("total shared wavefronts", tot)
for w, n, idl, s in items[:top]:
    print(f"  wavefronts={w:10d} exec={n:9d} per_exec={w / max(n, 1):5.2f} ideal_per_exec={idl / max(n, 1):5.2f}  {s[:90]}")
# aggregate by opcode and wavefronts-per-execution bucket
import collections
agg = collections.defaultdict(lambda: [0, 0, 0])
for w, n, idl, s in items:
    parts = s.split()
    op = parts[1] if parts[0].startswith('@') else parts[0]
    key = (op, round(w / max(n, 1)))
    agg[key][0] += 1; agg[key][1] += w; agg[key][2] += n
print("opcode, wavefronts/exec bucket: static instrs, wavefronts, executions")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k[0]:28s} ~{k[1]:2d}/exec  n_instr={v[0]:4d} wavefronts={v[1]:11d} ({100.0 * v[1] / tot:4.1f}%) exec={v[2]}")
