"""Per-kernel DRAM traffic of ONE generator forward from an ncu metrics CSV:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/traffic.csv python tools/ncu_target.py 16 861
  python tools/ncu_traffic.py gpurun_out/traffic.csv profiles/rNN_traffic          -> .txt (per kernel) + .json (read by bench.py)
"""
import collections, csv, json, sys
U = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
per = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(hdr, r))
    per.setdefault((int(d["ID"]), d["Kernel Name"].split("(")[0].replace("void ", "").replace("bvg::", "")), {})[d["Metric Name"]] = \
        float(d["Metric Value"].replace(",", "")) * U.get(d["Metric Unit"], 1)
agg = collections.OrderedDict()
cats = {"conv_tcgen05": [0, 0.0, 0.0, 0.0], "activation": [0, 0.0, 0.0, 0.0], "other": [0, 0.0, 0.0, 0.0]}
for (_, name), m in per.items():
    if name.startswith("pack_") or name.startswith("at::"):
        continue                                     # weight packing at load time / the target script's own checksum
    cat = "conv_tcgen05" if name.startswith("conv_umma") else "activation" if name.startswith("act1d") else "other"
    for a in (agg.setdefault(name, [0, 0.0, 0.0, 0.0]), cats[cat]):
        a[0] += 1; a[1] += m["dram__bytes_read.sum"]; a[2] += m["dram__bytes_write.sum"]; a[3] += m["gpu__time_duration.sum"]
with open(sys.argv[2] + ".txt", "w") as f:
    f.write("# DRAM traffic per kernel of one forward (16 x 861 frames, bf16 mode), ncu dram__bytes_read/write.sum, cold caches, serialised launches\n")
    for name, a in agg.items():
        f.write("%-66s n=%3d  read %9.1f MB  write %9.1f MB  %9.1f us\n" % (name[:66], a[0], a[1] / 1e6, a[2] / 1e6, a[3] / 1e3))
    for c, a in cats.items():
        f.write("== %-14s launches %3d  traffic %.2f GB (%.1f MB per launch)  %.2f ms\n" % (c, a[0], (a[1] + a[2]) / 1e9, (a[1] + a[2]) / 1e6 / max(a[0], 1), a[3] / 1e6))
json.dump({c: {"launches": a[0], "dram_bytes_read": a[1], "dram_bytes_write": a[2], "dram_bytes_per_launch": (a[1] + a[2]) / max(a[0], 1),
               "ncu_time_ms": a[3] / 1e6} for c, a in cats.items()} | {"workload": "16 x 861 frames, bf16", "source": sys.argv[1]},
          open(sys.argv[2] + ".json", "w"), indent=1)
print(open(sys.argv[2] + ".txt").read())
