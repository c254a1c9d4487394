"""List the loops (backward branches) of one kernel in a cuobjdump -sass dump with their instruction mix."""
import re, sys
from collections import Counter
txt = open(sys.argv[1]).read(); pat = sys.argv[2]
for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function :|\Z)", txt, re.S):
    if pat not in m.group(1): continue
    ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+((?:@!?U?P\d+\s+)?)([A-Z0-9_]+)([^;]*);", m.group(2))
    addrs = [int(a, 16) for a, _, _, _ in ins]
    print(m.group(1), len(ins), "instructions")
    for i, (a, pred, op, rest) in enumerate(ins):
        if op == "BRA":
            t = re.search(r"0x([0-9a-f]+)", rest)
            if t and int(t.group(1), 16) < int(a, 16):
                lo = int(t.group(1), 16)
                body = [o for (aa, _, o, _) in ins if lo <= int(aa, 16) <= int(a, 16)]
                c = Counter(body)
                print("  loop 0x%x..0x%s: %d instr:" % (lo, a, len(body)), dict(c.most_common(14)))
