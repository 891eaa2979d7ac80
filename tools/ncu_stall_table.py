"""Development aid: source-level (SASS) stall table of the CMUX loop of a PBS kernel from an ncu report captured with
--set full --import-source on.  Bins the loop body (the instructions executed lwe_dimension times) into 128-instruction
windows and prints, per window, the share of the loop's warp-state samples, the instruction mix and the top stall reasons.
usage: ncu_stall_table.py report.ncu-rep [kernel-regex]"""
import csv, io, subprocess, sys, collections, re

rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "pbs_kernel3"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
data = [r for r in rows[2:] if len(r) >= len(h) - 2 and r[0].startswith("0x")]
ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
ie = [int(r[ix["Instructions Executed"]]) for r in data]
big = collections.Counter(v for v in ie if v > 0).most_common(1)[0][0]
idx = [i for i, v in enumerate(ie) if v == big]
l0, l1 = idx[0], idx[-1] + 1
tot = sum(int(r[ix["# Samples"]]) for r in data[l0:l1])
print(f"{rows[0][1] if len(rows[0]) > 1 else kern}: loop body = SASS instructions {l0}..{l1} ({l1 - l0} instructions, each executed {big} times), "
      f"{tot} warp-state samples in the loop")
print("window  samples%  fp64  ldst  alu+fma  other | top warp states (% of the window's samples)")
def cls(s):
    op = s.split()[0] if not s.lstrip().startswith("@") else s.split()[1]
    op = op.split(".")[0]
    if op in ("DFMA", "DADD", "DMUL"): return "fp64"
    if op in ("LDS", "STS", "LDTM", "STTM", "LDG", "STG", "LDL", "STL", "UBLKCP", "ATOMS"): return "ldst"
    if op in ("IADD3", "LOP3", "SHF", "SEL", "ISETP", "VIADD", "IMAD", "LEA", "PRMT", "MOV"): return "int"
    return "other"
allt = collections.Counter()
for b in range(l0, l1, 128):
    hi = min(b + 128, l1)
    t = collections.Counter(); mix = collections.Counter()
    for r in data[b:hi]:
        for s in stalls: t[s[6:]] += int(r[ix[s]])
        mix[cls(r[ix["Source"]].strip())] += 1
    T = sum(t.values()) or 1
    allt.update(t)
    print(f"{b - l0:5d}   {100 * T / tot:6.2f}   {mix['fp64']:4d}  {mix['ldst']:4d}  {mix['int']:6d}  {mix['other']:5d} | " +
          ", ".join(f"{k} {100 * v / T:.0f}" for k, v in t.most_common(4)))
T = sum(allt.values())
print("whole loop: " + ", ".join(f"{k} {100 * v / T:.1f}%" for k, v in allt.most_common(10)))
print("\ntop 25 instructions by samples:")
top = sorted(range(l0, l1), key=lambda i: -int(data[i][ix["# Samples"]]))[:25]
for i in top:
    r = data[i]
    ss = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"  +{i - l0:5d}  {100 * int(r[ix['# Samples']]) / tot:5.2f}%  {r[ix['Source']].strip()[:70]:70s} {ss[0][1]} {ss[0][0]}, {ss[1][1]} {ss[1][0]}")
