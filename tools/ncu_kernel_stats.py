"""Per-kernel numbers bench.py's roofline quotes, from `ncu --set full` captures of the bench's own launches:
   python tools/ncu_kernel_stats.py <bench.json> <name> tag=path.ncu-rep [tag=path.ncu-rep ...]  > profiles/ndt_eval_traffic.json
tag in {exact_direct7, fast_direct7, exact_pca_direct1, fast_pca_direct1}.  Of the launches in a capture the longest one is taken (captures
of a whole align also contain launches that find nothing left to do).  algorithmic_bytes_per_launch comes from the bench line of the
same GPU visit, so that bench.py can scale the instruction count when it runs another batch size."""
import csv, json, subprocess, sys
bench = json.load(open(sys.argv[1])); name = sys.argv[2]
alg = {"exact_direct7": bench.get("roofline", {}), "fast_direct7": bench.get("modes", {}).get("tolerance", {}).get("roofline", {}),
       "exact_pca_direct1": bench.get("configs", {}).get("pca_direct1", {}).get("exact", {}).get("roofline", {}),
       "fast_pca_direct1": bench.get("configs", {}).get("pca_direct1", {}).get("tolerance", {}).get("roofline", {})}
mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "inst": 1, "usecond": 1, "us": 1, "msecond": 1e3, "ms": 1e3, "nsecond": 1e-3, "ns": 1e-3, "%": 1, "": 1}
out = {}
for arg in sys.argv[3:]:
    tag, path = arg.split("=", 1)
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    def val(r, k):
        i = h.index(k); return float(r[i].replace(",", "")) * mul.get(units[i], 1)
    best = max(rows[2:], key=lambda r: val(r, "gpu__time_duration.sum"))
    out[tag] = {"kernel": best[h.index("Kernel Name")], "grid_size": best[h.index("launch__grid_size")],
                "duration_us_under_ncu": val(best, "gpu__time_duration.sum"),
                "dram_bytes_per_launch": val(best, "dram__bytes_read.sum") + val(best, "dram__bytes_write.sum"),
                "warp_inst_per_launch": val(best, "smsp__inst_executed.sum"),
                "issue_active_pct": val(best, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "xu_pct": val(best, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                "fma_pct": val(best, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                "fp64_pct": val(best, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                "lsu_pct": val(best, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                "registers_per_thread": val(best, "launch__registers_per_thread"),
                "warps_active_pct": val(best, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "algorithmic_bytes_per_launch": alg.get(tag, {}).get("algorithmic_bytes_per_launch"),
                "source": "profiles/%s/%s (ncu --set full --clock-control none, longest captured launch)" % (name, path.split("/")[-1])}
print(json.dumps(out, indent=1))
