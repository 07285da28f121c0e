"""Summarise every launch of an `ncu --set full` report into one JSON kept under profiles/.
usage: python tools/ncu_summary_all.py gpurun_out/r02_prof.ncu-rep profiles/out.json "note ..." """
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg", "smsp__warps_eligible.avg.per_cycle_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALL = "smsp__average_warps_issue_stalled_"

rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, u = rows[0], rows[1]
col = {n: i for i, n in enumerate(h)}
launches = []
for v in rows[2:]:
    top = {"kernel": v[col["Kernel Name"]]}
    for k in KEYS:
        if k in col:
            top[k] = f"{v[col[k]]} {u[col[k]]}".strip()
    stalls = {}
    for n, i in col.items():
        if n.startswith(STALL) and n.endswith("_per_issue_active.ratio"):
            x = float(v[i])
            if x >= 0.1:
                stalls[n[len(STALL):-len("_per_issue_active.ratio")]] = round(x, 2)
    top["warp_stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    launches.append(top)
json.dump({"note": note, "launches": launches}, open(out, "w"), indent=1)
print(len(launches), "launches ->", out)
