"""Turns the raw outputs of a measurement session (gpurun_out/) into the tracked summaries under profiles/.

    python tools/make_profiles.py r02 bench.json ref.json launches.csv full1.ncu-rep [full2.ncu-rep ...]
"""
import collections, csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, bench_json, ref_json, launches_csv = sys.argv[1:5]
reps = sys.argv[5:]
traffic_reps = [r for r in reps if "cliff" not in os.path.basename(r)]  # (the cliff capture is a different layer)
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

def last_json_line(path):
    return json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])

if os.path.exists(bench_json):
    json.dump(last_json_line(bench_json), open(os.path.join(P, f"{tag}_bench_n1.json"), "w"), indent=1)
if os.path.exists(ref_json):
    json.dump(last_json_line(ref_json), open(os.path.join(P, f"{tag}_bench_reference_arm.json"), "w"), indent=1)

def short(name):
    return re.sub(r"\(.*", "", name).replace("void ", "").strip()

# ---- launch list ------------------------------------------------------------------------------------------------
if os.path.exists(launches_csv):
    rows = [r for r in csv.reader(l for l in open(launches_csv) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    with open(os.path.join(P, f"{tag}_launches.csv"), "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 1 --warmup 1 --batch 4 --resident 12 --no-e2e --no-cpu-baseline --no-extras (the headline step only)\n")
        fh.write("id,kernel,duration_us\n")
        for r in rows:
            v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1.0)
            fh.write(f"{r[0]},{short(r[kn])},{v:.3f}\n")
            agg.setdefault(short(r[kn]), []).append(v)
    setup = lambda k: "synth" in k or k.startswith("at::")   # stack rasteriser / torch fills: not part of a step
    tot = sum(sum(v) for k, v in agg.items() if not setup(k))
    with open(os.path.join(P, f"{tag}_launches_summary.csv"), "w") as fh:
        fh.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache serialised launches: compare SHARES)\n")
        fh.write("kernel,launches,mean_us,share_of_all_listed_kernels\n")
        for k, v in agg.items():
            fh.write(f"{k.replace(',', ';')},{len(v)},{sum(v)/len(v):.2f},{(sum(v)/tot if not setup(k) else 0):.4f}\n")

# ---- ncu --set full summaries --------------------------------------------------------------------------------------
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
out, traffic = [], collections.defaultdict(list)
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    if len(r) < 3:
        continue
    h, units, data = r[0], r[1], r[2:]
    for row in data:
        d = dict(zip(h, row)); u = dict(zip(h, units))
        e = {"capture": os.path.basename(rep), "Kernel Name": d["Kernel Name"]}
        for k in keep:
            if k in d:
                e[k] = f"{d[k]} {u[k]}".strip()
        st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(d[k].replace(",", "")) for k in h
              if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and d[k]}
        tot = sum(st.values()) or 1.0
        e["stall_shares"] = {k: round(v / tot, 3) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]}
        out.append(e)
        def mb(k):
            v = float(d[k].replace(",", "")); return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[u[k]]
        if rep in traffic_reps:
            traffic[short(d["Kernel Name"])].append(mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"))
if out:
    json.dump(out, open(os.path.join(P, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
    per = {k: round(sum(v) / len(v), -4) for k, v in traffic.items()}
    t = {"per_kernel": per,
         # the bench's stage names: bytes per layer of all the launches the stage makes
         "layer_fused": sum(v for k, v in per.items() if (k.startswith("k_layer2") and "<4," not in k) or (k.startswith("k_layer<") and k.endswith(", 1>"))),
         "layer_patch_lut_only": sum(v for k, v in per.items() if k.startswith("k_layer<") and k.endswith(", 0>")),
         "pack_flags_lut_only": sum(v for k, v in per.items() if k.startswith("k_pack_tma<1>")),
         "pack_flags": sum(v for k, v in per.items() if k.startswith("k_pack_tma<0>") or k.startswith("k_pack_tma<(bool)0>")),
         "_source": f"profiles/{tag}_ncu_full_summary.json (ncu --set full --clock-control none), bytes per launch = dram__bytes_read.sum + "
                    "dram__bytes_write.sum, mean over the captured launches; layer_fused = k_layer2 + the list-mode k_layer launch"}
    json.dump(t, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    print(json.dumps(t, indent=1))
