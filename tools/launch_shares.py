"""Per-kernel share of the device time in an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python tools/launch_shares.py profiles/r02_final_launches.csv [first_launch last_launch] > profiles/r02_final_launch_shares.txt
(cold-cache, serialised per-launch times: the SHARES are what must agree with bench.py's live CUDA-event profile)."""
import csv, re, sys
from collections import OrderedDict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = next(r for r in rows if "Kernel Name" in r)
data = [r for r in rows if r is not hdr and r[0].strip().isdigit()]
iN, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
agg = OrderedDict()
for r in data[lo:hi]:
    name = re.sub(r"<.*", "", r[iN].replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", ""))
    us = float(r[iV].replace(",", "")) * scale.get(r[iU], 1.0)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]} launches [{lo}, {hi}): {hi - lo} launches, {tot:.1f} us of kernel time")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{100 * us / tot:6.2f} %  {us:9.1f} us  {n:4d} x  {k}")
