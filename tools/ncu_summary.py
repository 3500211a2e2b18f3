"""Key metrics of an .ncu-rep (first kernel): python tools/ncu_summary.py X.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    for vals in rows[2:]:   # every captured launch of the report
        d = dict(zip(hdr, vals))
        print("==", path, "|", d.get("Kernel Name", "")[:70])
        for w in WANT:
            if w in d:
                print(f"  {w:70s} {d[w]} {u[w]}")
        stalls = sorted(((float(v), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")),
                        reverse=True)
        print("  stalls/issue:", ", ".join(f"{k[34:-23]}={v:.2f}" for v, k in stalls[:7]))
