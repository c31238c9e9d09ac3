"""ptxas-level register report for one kernel of libdib_b200.so: static LDL/STL counts per 500-instruction bucket next to
the landmarks (UTCHMMA = MMA issue loop, LDTM = epilogues, STG = flush).  Usage: python tools/spill_report.py <mangled-substr>"""
import re, subprocess, sys
so = "distributed-information-bottleneck.github.io_b200/libdib_b200.so"
funs = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", funs)
want = sys.argv[1]
for b in blocks[1:]:
    name = b.split("\n", 1)[0]
    if want not in name:
        continue
    ops = []
    for l in b.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if m:
            ops.append(m.group(2))
    print(name[:110], "instructions", len(ops), "LDL", sum(o.startswith("LDL") for o in ops), "STL", sum(o.startswith("STL") for o in ops))
    B = 500
    for s in range(0, len(ops), B):
        seg = ops[s:s + B]
        c = lambda p: sum(o.startswith(p) for o in seg)
        print(f"  {s:5d}: LDL {c('LDL'):3d} STL {c('STL'):3d} | UTCHMMA {c('UTCHMMA'):3d} LDTM {c('LDTM'):2d} MUFU {c('MUFU'):3d} SYNCS {c('SYNCS'):3d} STS {c('STS'):3d} STG {c('STG'):3d} LDG {c('LDG'):3d}")
