"""Summarises an .ncu-rep (ncu --set full) as JSON for profiles/: duration, FP64 pipe, issue slots, occupancy,
DRAM bytes, L2 hit rate, shared-memory conflicts, atomics, top stall reasons.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.json ["note"]"""
import csv, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed_op_global_red.sum",
        "lts__t_sectors_srcunit_tex_op_red.sum", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
res = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    k = {"kernel": d.get("Kernel Name"), "metrics": {}, "stalls_per_issue": {}}
    for key in KEYS:
        if d.get(key) not in (None, ""):
            try:
                k["metrics"][key] = {"value": float(d[key].replace(",", "")), "unit": units[hdr.index(key)]}
            except ValueError:
                pass
    for key in hdr:
        if "issue_stalled" in key and key.endswith("per_issue_active.ratio"):
            try:
                v = float(d[key])
            except ValueError:
                continue
            if v >= 0.05:
                k["stalls_per_issue"][key.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = v
    m = k["metrics"]
    if "dram__bytes_read.sum" in m and "dram__bytes_write.sum" in m:
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        k["dram_bytes_per_launch"] = sum(m[q]["value"] * scale.get(m[q]["unit"], 1) for q in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    res.append(k)
json.dump({"capture": rep.split("/")[-1], "command": "ncu --set full --clock-control none --import-source on", "note": note,
           "kernels": res}, open(out, "w"), indent=1)
print("wrote", out, [(k["kernel"][:50], k["metrics"].get("gpu__time_duration.sum")) for k in res])
