"""Top stall sites of one kernel in an .ncu-rep (source page, SASS view).
usage: python tools/ncu_hot.py rep [kernel-index (0-based)] [top-N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
# the csv holds one table per kernel, each introduced by a "Kernel Name" line
tables, cur = [], None
for ln in out.splitlines():
    if ln.startswith('"Kernel Name"'):
        cur = {"name": ln, "lines": []}; tables.append(cur)
    elif cur is not None:
        cur["lines"].append(ln)
t = tables[2 * kidx] if len(tables) > kidx and len(tables) % 2 == 0 and tables[0]["name"] == tables[1]["name"] else tables[kidx]
print(t["name"][:160])
rows = list(csv.reader(io.StringIO("\n".join(t["lines"]))))
hdr = rows[0]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
body = [r for r in rows[1:] if len(r) == len(hdr)]
tot = sum(int(r[isamp] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
order = sorted(range(len(body)), key=lambda i: -int(body[i][isamp] or 0))[:topn]
for i in sorted(order):
    r = body[i]
    ctx = body[i - 1][isrc].strip()[:50] if i else ""
    print("%5d %6.2f%% ex=%9s  %-60s | prev: %s" % (i, 100.0 * int(r[isamp] or 0) / max(tot, 1), r[iex], r[isrc].strip()[:60], ctx))
