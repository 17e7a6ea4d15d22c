"""Copies the judged summaries of one GPU round from gpurun_out/<tag>/ into profiles/<name>/ (tracked).
   python tools/make_profiles.py r01 round1_a"""
import collections, csv, json, os, shutil, subprocess, sys
tag, name = sys.argv[1], sys.argv[2]
src = os.path.join("gpurun_out", tag); dst = os.path.join("profiles", name); os.makedirs(dst, exist_ok=True)
for f in ("bench.json", "bench_pca.json", "bench_ref.json", "pytest_gpu.log", "smoke.log", "gpu_first.log", "pgo_perf.log", "aux_perf.log", "chol_solve.log", "gpu.txt", "nproc.txt", "breakdown.log",
          "launches.csv", "shard.log", "bench_n2.json"):
    if os.path.exists(os.path.join(src, f)): shutil.copy(os.path.join(src, f), os.path.join(dst, f))
# launch list -> per-kernel shares
lc = os.path.join(src, "launches.csv")
if os.path.exists(lc):
    rows = [r for r in csv.reader(open(lc)) if len(r) > 5]
    hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try: v = float(r[vi].replace(',', ''))
        except ValueError: continue
        k = r[ki].split('(')[0]; agg[k][0] += 1; agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(dst, "launch_shares.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-40s launches %5d  total %10.1f us  avg %8.2f us  share %5.1f%%\n" % (k, v[0], v[1] / 1e3, v[1] / 1e3 / v[0], 100 * v[1] / tot))
# full captures -> raw metric summary + per-source-line instruction table
for rep in sorted(f for f in os.listdir(src) if f.endswith(".ncu-rep")):
    base = rep[:-8]
    raw = subprocess.run(["ncu", "-i", os.path.join(src, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open("/tmp/_raw.csv", "w").write(raw)
    out = subprocess.run([sys.executable, "tools/ncu_summary.py", "/tmp/_raw.csv"], capture_output=True, text=True).stdout
    open(os.path.join(dst, base + "_ncu_summary.txt"), "w").write(out)
    # the capture holds two launches of an align; the working one is the longest (the other may find nothing left to do)
    rows = list(csv.reader(raw.splitlines()))
    longest = 1
    if len(rows) > 2 and 'gpu__time_duration.sum' in rows[0]:
        di = rows[0].index('gpu__time_duration.sum')
        durs = [float(r[di].replace(',', '')) for r in rows[2:]]
        longest = 1 + durs.index(max(durs))
    sp = subprocess.run(["ncu", "-i", os.path.join(src, rep), "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::%d" % longest],
                        capture_output=True, text=True).stdout
    open("/tmp/_src.csv", "w").write(sp)
    out = subprocess.run([sys.executable, "tools/ncu_lines.py", "/tmp/_src.csv", "40"], capture_output=True, text=True).stdout
    open(os.path.join(dst, base + "_ncu_source_lines.txt"), "w").write(out)
    # dram traffic of the working launch (per-capture record; profiles/ndt_eval_traffic.json, which bench.py quotes, is written by
    # tools/ncu_kernel_stats.py from the same captures)
    if len(rows) > 2:
        h = rows[0]
        try:
            r = rows[1 + longest]
            rd = float(r[h.index('dram__bytes_read.sum')]); wr = float(r[h.index('dram__bytes_write.sum')])
            ur, uw = rows[1][h.index('dram__bytes_read.sum')], rows[1][h.index('dram__bytes_write.sum')]
            mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            json.dump({"kernel": r[h.index('Kernel Name')], "dram_bytes_per_launch": rd * mul[ur] + wr * mul[uw],
                       "duration_us_under_ncu": float(r[h.index('gpu__time_duration.sum')]), "grid_size": r[h.index('launch__grid_size')], "source": "%s/%s (ncu --set full, longest captured launch)" % (name, rep)},
                      open(os.path.join(dst, base + "_traffic.json"), "w"))
        except (ValueError, KeyError): pass
print(sorted(os.listdir(dst)))
