"""Key counters + stall shares of every kernel in an .ncu-rep (ncu --set full capture): python tools/ncu_summary.py rep [out.json]"""
import csv, json, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, units, data = r[0], r[1], r[2:]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum"]
out = []
for row in data:
    d = dict(zip(h, row)); u = dict(zip(h, units))
    e = {"Kernel Name": d["Kernel Name"][:60]}
    for k in keep:
        if k in d:
            e[k] = f"{d[k]} {u[k]}".strip()
    st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(d[k].replace(",", "")) for k in h
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and d[k]}
    tot = sum(st.values()) or 1.0
    e["stall_shares"] = {k: round(v / tot, 3) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]}
    out.append(e)
print(json.dumps(out, indent=1))
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], "w"), indent=1)
