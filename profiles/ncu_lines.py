"""Samples / instructions per CUDA source line of an .ncu-rep (needs -lineinfo + --import-source on):
    python profiles/ncu_lines.py report.ncu-rep [min_pct]"""
import csv, os, subprocess, sys
rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, h, items = "", None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        h = r; sm, ei = h.index("# Samples"), h.index("Instructions Executed")
        wv = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else None
        continue
    if h and r[0].isdigit():
        try:
            items.append((cur_file, int(r[0]), r[1].strip()[:100], int(r[sm] or 0), int(r[ei] or 0), int(r[wv] or 0) if wv is not None else 0))
        except ValueError:
            pass
ts, tn, tw = (sum(x[i] for x in items) or 1 for i in (3, 4, 5))
print(f"total samples {ts}, warp instructions {tn}, shared wavefronts {tw}")
for f, ln, src, s, n, w in items:
    if max(100.0 * s / ts, 100.0 * n / tn, 100.0 * w / tw) >= min_pct:
        print(f"  {100.0 * s / ts:5.1f}% smp {100.0 * n / tn:5.1f}% ins {100.0 * w / tw:5.1f}% wav  {f}:{ln}  {src}")
