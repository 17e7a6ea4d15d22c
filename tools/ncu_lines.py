"""Instruction counts per CUDA source line from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
cur = None; agg = collections.Counter(); samp = collections.Counter(); src = {}; ie = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) > 3 and r[0] == "Line No": ie = r.index('Instructions Executed'); isamp = r.index('# Samples'); continue
    if ie is None or len(r) <= ie or not r[0] or r[2] != '-': continue
    try: e = int(r[ie]); s = int(r[isamp])
    except ValueError: continue
    k = (cur, int(r[0])); agg[k] += e; samp[k] += s; src[k] = r[1].strip()[:120]
tot = sum(agg.values()); ts = max(1, sum(samp.values()))
print("total warp instructions", tot)
for k, v in agg.most_common(top): print("%5.1f%% inst %5.1f%% samp  %-22s %4d  %s" % (100 * v / tot, 100 * samp[k] / ts, k[0], k[1], src[k]))
