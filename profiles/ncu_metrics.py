"""Print selected metrics of an .ncu-rep: python profiles/ncu_metrics.py file.ncu-rep [extra regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
WANT = [r"^gpu__time_duration\.sum$", r"^dram__bytes_(read|write)\.sum$", r"^gpu__dram_throughput\.avg\.pct", r"^lts__t_bytes\.sum$",
        r"^lts__throughput\.avg\.pct", r"^l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed", r"^sm__throughput\.avg\.pct",
        r"^sm__pipe_tensor.*cycles_active.*(avg|pct)", r"^sm__cycles_elapsed\.max$", r"^launch__(grid_size|registers_per_thread|cluster_size|shared_mem_per_block_dynamic)$",
        r"^sm__warps_active\.avg\.pct", r"^smsp__inst_executed\.sum$", r"^smsp__average_warps?_issue_stalled_.*_per_issue_active\.ratio$",
        r"^smsp__issue_active\.avg\.pct", r"^lts__t_sector_hit_rate\.pct$", r"^l1tex__t_sector_hit_rate\.pct$", r"^lts__t_sectors_srcunit_tex_op_(read|write)\.sum$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared.*sum$", r"^smsp__warp_issue_stalled_.*_per_warp_active\.pct$"]
if extra:
    WANT.append(extra)
for k in range(2, len(rows)):
    vals = rows[k]
    name = dict(zip(hdr, vals)).get("Kernel Name", "")[:100]
    print("==", name)
    for h, u, v in zip(hdr, units, vals):
        if any(re.search(w, h) for w in WANT):
            print(f"  {h:90s} {u:10s} {v}")
