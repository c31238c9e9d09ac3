"""Turn an .ncu-rep into a short text summary (per captured kernel: duration, DRAM bytes, pipe utilisation,
occupancy limits, top stall reasons).  Usage: python tools/summarize_ncu.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum"]
print(f"# {rep}")
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print(f"\n## {name[:150]}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:75s} {r[i]:>18s} {units[i]}")
    stalls = [(float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, h in enumerate(hdr)
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and r[i].replace(".", "").isdigit()]
    tot = sum(s for s, _ in stalls) or 1
    print("  stall samples: " + ", ".join(f"{n} {100 * s / tot:.0f}%" for s, n in sorted(stalls, reverse=True)[:7]))
