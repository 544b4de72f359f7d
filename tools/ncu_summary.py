"""Markdown summary of an .ncu-rep (run where ncu is installed; no GPU needed).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep "title / context" > profiles/r01_x.md"""
import csv, subprocess, sys, io

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of peak warps)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "warp execution efficiency (active threads / instruction, of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue-slot utilisation (%)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe utilisation (%)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__bytes_read.sum.per_second", "DRAM read rate"),
    ("lts__t_bytes.sum.per_second", "L2 throughput"), ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate (%)"),
    ("lts__t_sectors_op_red.sum", "L2 sectors, reductions (red.global.add)"),
    ("lts__t_sectors_op_atom.sum", "L2 sectors, atomics with return"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "cycles between issues of one warp"),
]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
print(f"# ncu summary: {title}\n\nreport: `{rep}`\n")
for r in rows[2:]:
    kname = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print(f"## kernel `{kname}`\n\n| metric | value |\n|---|---|")
    for key, label in WANT:
        if key in hdr:
            i = hdr.index(key)
            if r[i] != "":
                print(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
    print("\n| warp stall reason (per issue) | value |\n|---|---|")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i] != "":
            st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    for v, n in sorted(st, reverse=True)[:8]:
        print(f"| {n} | {v:.3f} |")
    print()
