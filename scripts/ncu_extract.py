"""Extract the per-launch metrics that DESIGN.md / bench.py quote from an ncu report (run here, no GPU needed).
   python scripts/ncu_extract.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_<what>.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe % active"),
        ("smsp__inst_executed.sum", "warp instructions"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "registers"), ("launch__occupancy_limit_registers", "occupancy limit (regs)"),
        ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %")]
print("# %s (ncu --set full --clock-control none); one block per captured launch" % rep)
for r in rows[2:]:
    print("\n%s  grid %s block %s" % (r[ix["Kernel Name"]][:110], r[ix["Grid Size"]], r[ix["Block Size"]]))
    for key, label in want:
        if key in ix:
            print("    %-28s %14s %s" % (label, r[ix[key]][:14], units[ix[key]]))
    try:
        tot = float(r[ix["dram__bytes_read.sum"]]) + float(r[ix["dram__bytes_write.sum"]])
        u = units[ix["dram__bytes_read.sum"]]
        t = float(r[ix["gpu__time_duration.sum"]])
        tu = units[ix["gpu__time_duration.sum"]]
        print("    %-28s %14.4f %s  (read+write)" % ("dram traffic", tot, u))
        if u == "Gbyte" and tu in ("ms", "msecond"):
            print("    %-28s %14.1f GB/s" % ("dram bandwidth", tot / (t * 1e-3)))
    except Exception:
        pass
