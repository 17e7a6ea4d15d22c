"""Launch list of the last direct solve in an ncu CSV (ncu --metrics gpu__time_duration.sum,launch__grid_size ... python tools/chol_profile.py 100 50 1)."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]; ki = h.index('Kernel Name'); mi = h.index('Metric Name'); vi = h.index('Metric Value'); ii = h.index('ID')
cur, order = {}, []
for r in rows[1:]:
    k = (r[ii], r[ki].split('(')[0][:34])
    if k not in cur: order.append(k)
    cur.setdefault(k, {})[r[mi]] = r[vi]
idx = [i for i, k in enumerate(order) if 'scatter' in k[1]]
fin = [i for i, k in enumerate(order) if 'finish' in k[1]]
a, b = idx[-1], max(i for i in fin if i > idx[-1]) + 1
tot = f = bk = 0
for k in order[a:b]:
    d = float(cur[k]['gpu__time_duration.sum']) / 1e3
    print("%-36s grid %-6s %8.1f us" % (k[1], cur[k]['launch__grid_size'], d))
    tot += d; f += d if 'front' in k[1] else 0; bk += d if 'backward' in k[1] else 0
print("# total %.1f us: forward (fronts) %.1f us, backward %.1f us" % (tot, f, bk))
