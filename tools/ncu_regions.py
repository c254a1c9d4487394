"""Where the warps of a kernel spend their time, from the source page of an .ncu-rep (needs -lineinfo / --import-source on):
stall-sample totals by reason, the SASS regions (blocks of 250 instructions) with their instruction / sample shares and
top stall reasons, and the 15 individual instructions with the most samples.
  python tools/ncu_regions.py rep.ncu-rep >> profiles/x.txt"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
ie, isamp = h.index("Instructions Executed"), h.index("# Samples")
reasons = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
ci = {c: h.index(c) for c in reasons}
tot_i = sum(int(r[ie]) for r in body); tot_s = sum(int(r[isamp]) for r in body)
print("== %s: %d SASS instructions, %d executed (thread-level sum), %d warp samples" % (rep, len(body), tot_i, tot_s))
tot = {c: sum(int(r[ci[c]]) for r in body) for c in reasons}
print("stall samples by reason: " + ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / tot_s) for c, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > tot_s / 200))
def opname(t):
    p = t.split()
    return p[1] if p and p[0].startswith("@") and len(p) > 1 else (p[0] if p else "?")
print("regions (250 SASS instructions each): first index, share of executed instructions, share of samples, top stall reasons, top opcodes")
for s in range(0, len(body), 250):
    seg = body[s:s + 250]
    n = sum(int(r[ie]) for r in seg); sm = sum(int(r[isamp]) for r in seg)
    if sm < tot_s / 100:
        continue
    st = sorted(((c[6:], sum(int(r[ci[c]]) for r in seg)) for c in reasons), key=lambda kv: -kv[1])[:4]
    ops = collections.Counter()
    for r in seg:
        ops[opname(r[1])] += int(r[ie])
    print("  %5d  inst %5.1f%%  samples %5.1f%%  %s  %s" % (s, 100.0 * n / max(tot_i, 1), 100.0 * sm / tot_s,
          " ".join("%s=%d" % kv for kv in st), ",".join(o for o, _ in ops.most_common(3))))
print("instructions with the most samples:")
for i in sorted(sorted(range(len(body)), key=lambda i: -int(body[i][isamp]))[:15]):
    r = body[i]
    st = sorted(((c[6:], int(r[ci[c]])) for c in reasons), key=lambda kv: -kv[1])[:2]
    print("  %5d  executed %9s  samples %5s  %s  %s" % (i, r[ie], r[isamp], " ".join("%s=%d" % kv for kv in st), r[1].strip()[:70]))
