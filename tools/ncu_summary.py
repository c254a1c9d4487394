"""Compact text summary of an .ncu-rep (raw page): python tools/ncu_summary.py rep [rep...] > profiles/x.txt"""
import csv, io, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__cycles_active.avg"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("== %s" % rep)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("kernel: %s" % d.get("Kernel Name", "?")[:110])
        for w in WANT:
            if w in d:
                print("   %-70s %s %s" % (w, d[w], units[hdr.index(w)]))
        rd = float(d.get("dram__bytes_read.sum", 0) or 0); wr = float(d.get("dram__bytes_write.sum", 0) or 0)
        print("   dram traffic (read+write)                                              %.3f %s" % (rd + wr, units[hdr.index("dram__bytes_read.sum")]))
