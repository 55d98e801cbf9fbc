""".ncu-rep (ncu --set full capture) -> short text summary of the metrics the roofline discussion uses.
    python scripts/ncu_summary.py gpurun_out/<tag>/prof_<kernel>.ncu-rep > profiles/<round>_<kernel>_ncu.txt"""
import csv, io, subprocess, sys

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "sm__cycles_elapsed.max",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor",
    "smsp__average_warp", "smsp__warp_issue_stalled", "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu", "sm__inst_executed_pipe_alu",
    "sm__inst_executed_pipe_fma", "smsp__cycles_active.avg",
]
EXACT_SKIP = ("_realtime", "dmma", "imma", ".min.", ".max.pct", ".sum.pct", "per_second")

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("== %s" % name[:150])
    for i, h in enumerate(hdr):
        if any(w in h for w in WANT) and not any(s in h for s in EXACT_SKIP):
            v = r[i]
            if v in ("", "n/a"):
                continue
            if "warp_issue_stalled" in h and not h.endswith("_per_warp_active.pct"):
                continue
            print("   %-92s %s %s" % (h, v, units[i]))
    dr, dw = r[hdr.index("dram__bytes_read.sum")], r[hdr.index("dram__bytes_write.sum")]
    print("   traffic = dram read + write = %s %s + %s %s" % (dr, units[hdr.index("dram__bytes_read.sum")], dw, units[hdr.index("dram__bytes_write.sum")]))
