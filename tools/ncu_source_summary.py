"""Summarise `ncu --page source --csv` output: hottest SASS lines by stall samples and the opcode mix."""
import csv, sys
from collections import Counter
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = next(r for r in rows if "Source" in r and "Address" in r)
data = [r for r in rows if r is not hdr and r[0].startswith("0x")]
iS, iI, isrc = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Source")
num = lambda v: int(v) if v.strip().isdigit() else 0
tot, toti = sum(num(r[iS]) for r in data), sum(num(r[iI]) for r in data)
print("total samples", tot, "total warp instr", toti, "n sass", len(data))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in sorted(data, key=lambda r: -num(r[iS]))[:ntop]:
    print(r[iS].rjust(6), r[iI].rjust(9), r[isrc][:120])
c, cs = Counter(), Counter()
for r in data:
    toks = r[isrc].split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = op.split(".")[0]
    c[op] += num(r[iI]); cs[op] += num(r[iS])
print("--- opcode mix (warp instr, %, samples)")
for op, n in c.most_common(30):
    print(op.ljust(12), str(n).rjust(11), f"{100 * n / max(toti, 1):5.1f}%", str(cs[op]).rjust(7))
