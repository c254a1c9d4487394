"""Per-kernel totals and shares of an ncu launch list (--metrics gpu__time_duration.sum --csv):
  python tools/ncu_launches.py gpurun_out/launches.csv "<header comment>" > profiles/rNN_launches.txt"""
import collections, csv, sys
U = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
agg = collections.OrderedDict()
tot = 0.0; n = 0
for r in rows[1:]:
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("bvg::", "")
    us = float(d["Metric Value"].replace(",", "")) * U[d["Metric Unit"]]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; tot += us; n += 1
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("launches %d total us %.1f" % (n, tot))
cat = collections.Counter()
for name, (k, us) in agg.items():
    cat["tcgen05 convs" if name.startswith(("conv_umma", "amp_unit")) else "stand-alone activations" if name.startswith("act1d") else "other"] += us
print("share: " + ", ".join("%s %.1f %%" % (c, 100 * v / tot) for c, v in cat.most_common()))
for name, (k, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-70s n=%4d  %10.1f us  %5.1f%%" % (name[:70], k, us, 100 * us / tot))
