"""Rank source lines of an ncu source-page CSV (ncu -i rep --page source --csv --print-source sass,cuda)
by stall samples and executed instructions.   python tools/ncu_lines.py file.csv [points] [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
points = float(sys.argv[2]) if len(sys.argv) > 2 else 1e6
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out, cur = [], None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if len(r) > 8 and r[0].isdigit() and r[2] == '-':
        try: out.append((int(r[6]), int(r[7]), cur, int(r[0]), r[1].strip()[:110]))
        except ValueError: pass
tot = sum(o[1] for o in out); ts = sum(o[0] for o in out)
print("total warp-instructions %.3e (%.1fk per 32 points), samples %d" % (tot, tot / (points / 32) / 1e3, ts))
for o in sorted(out, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100 * o[0] / ts, 100 * o[1] / tot, o[2], o[3], o[4]))
byfile = collections.Counter()
for o in out: byfile[o[2]] += o[1]
print({k: "%.1f%%" % (100 * v / tot) for k, v in byfile.items()})
