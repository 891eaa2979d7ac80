"""Development aid: decode the static scheduling control bits (stall count, yield, barriers) of a kernel's
SASS (cuobjdump -sass output) and print per-region sums -- a static estimate of single-warp issue cycles.
usage: sass_sched.py file.sass [start_idx end_idx]"""
import re, sys
ins = []
lines = open(sys.argv[1]).read().split("\n")
i = 0
pat = re.compile(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/")
pat2 = re.compile(r"^\s+/\* (0x[0-9a-f]{16}) \*/")
while i < len(lines):
    m = pat.match(lines[i])
    if m and i + 1 < len(lines):
        m2 = pat2.match(lines[i + 1])
        if m2:
            hi = int(m2.group(1), 16)
            ctrl = (hi >> 41) & 0x1FFFFF
            ins.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=ctrl & 0xF, yld=(ctrl >> 4) & 1,
                            wr=(ctrl >> 5) & 7, rd=(ctrl >> 8) & 7, wait=(ctrl >> 11) & 0x3F))
            i += 2
            continue
    i += 1
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi_ = int(sys.argv[3]) if len(sys.argv) > 3 else len(ins)
verbose = len(sys.argv) > 4
tot = 0
from collections import Counter
c = Counter(); cs = Counter()
for k in range(lo, hi_):
    d = ins[k]
    op = d["text"].split()[0] if not d["text"].startswith("@") else d["text"].split()[1]
    op = op.split(".")[0]
    c[op] += 1; cs[op] += d["stall"]
    tot += d["stall"]
    if verbose:
        print(f'{k:5d} {d["addr"]:6x} st={d["stall"]:2d} y={d["yld"]} wr={d["wr"]} rd={d["rd"]} wait={d["wait"]:06b}  {d["text"]}')
print(f"instructions {hi_-lo}, sum of stall counts {tot}")
for op, n in c.most_common(25):
    print(f"  {op:10s} n={n:5d} stall_sum={cs[op]:6d} avg={cs[op]/n:.2f}")
