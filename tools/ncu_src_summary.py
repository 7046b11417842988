"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per source line.

    python tools/ncu_src_summary.py src.csv [top]

Prints, per (file, line): warp-instructions executed, stall samples, avg active threads.
"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])  # inst, samples, thread inst, text
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None: continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    def g(name):
        i = hdr.get(name)
        try: return float(r[i]) if i is not None and r[i] != "" else 0.0
        except ValueError: return 0.0
    a = agg[(cur, line)]
    a[0] += g("Instructions Executed"); a[1] += g("# Samples"); a[2] += g("Thread Instructions Executed")
    if not a[3]: a[3] = r[1].strip()[:90]
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp-inst %.3g  samples %d" % (ti, ts))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-18s %5d  inst %5.1f%%  samp %5.1f%%  thr %4.1f  %s" % (f, l, 100 * a[0] / ti, 100 * a[1] / ts, a[2] / a[0] if a[0] else 0, a[3]))

# phase table for tile_core.cuh line ranges (edit when the file moves)
if len(sys.argv) > 3:
    ranges = [tuple(x.split(":")) for x in sys.argv[3:]]  # name:file:lo:hi
    for name, f, lo, hi in ranges:
        lo, hi = int(lo), int(hi)
        i = sum(a[0] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi)
        s = sum(a[1] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi)
        t = sum(a[2] for (ff, l), a in agg.items() if ff == f and lo <= l <= hi)
        print("%-14s inst %5.1f%%  samples %5.1f%%  thr %4.1f" % (name, 100 * i / ti, 100 * s / ts, t / i if i else 0))
