"""Summarise one or more .ncu-rep files (ncu --set full) into the text format of profiles/*_ncu_full_summary.txt and
print the per-launch DRAM traffic (for profiles/traffic.json)."""
import csv, json, subprocess, sys
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
           "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct",
           "lts__t_sectors_srcunit_tex_op_atom.sum"]
traffic = {}
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        name = r[H.index("Kernel Name")]
        short = name.split("(")[0].split("::")[-1]
        if short in seen:
            continue
        seen.add(short)
        print("## %s" % name[:80])
        vals = {}
        for m in METRICS:
            if m in H:
                i = H.index(m)
                print("  %-72s %s %s" % (m, r[i], U[i]))
                vals[m] = (float(r[i].replace(",", "")), U[i])
        def to_bytes(m):
            v, u = vals.get(m, (0.0, "byte"))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        key = short.split("<")[0]
        if key == "fused_project_bwd_kernel" and short.rstrip().endswith(", 1>"):
            key = "fused_project_bwd_adam_kernel"  # the ADAM = true instantiation
        traffic[key] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        print()
print("TRAFFIC", json.dumps(traffic))
