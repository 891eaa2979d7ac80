"""Development aid: per-warp phase timeline of CTA 0 of pbs_kernel3 (B200TFHE_PBS_TIMELINE=<file>)."""
import sys
import numpy as np
names = sys.argv[2].split(",") if len(sys.argv) > 2 else ["A", "fwdFFT", "stsF", "bskwait", "ownmul", "bar1", "othmul", "bar2", "invFFT", "D"]
rows = [list(map(int, l.split())) for l in open(sys.argv[1])]
d = {}
for r in rows:
    d[(r[0], r[1])] = np.array(r[2:13], dtype=np.int64)
t00 = min(v[0] for v in d.values() if v[0] > 0)
print("step warp start  " + " ".join(f"{n:>7}" for n in names) + "   total")
for step in range(8):
    for w in range(8):
        v = d[(step, w)]
        if v[0] == 0:
            continue
        dur = np.diff(v)
        print(f"{step:4d} {w:4d} {v[0]-t00:6d} " + " ".join(f"{x:7d}" for x in dur) + f" {v[10]-v[0]:7d}")
    print()
