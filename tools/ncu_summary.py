"""Summarise an `ncu --page raw --csv` export per kernel: launches, time, DRAM bytes, warp instructions, pipe / issue
utilisation (time-weighted). usage: ncu_summary.py raw.csv images_per_launch [traffic.json] [launches.csv]
(prints a markdown table; optionally writes the per-image DRAM traffic json bench.py reads and a launch list)."""
import csv, json, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
imgs = float(sys.argv[2])
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
def f(r, n):
    try: return float(r[col[n]])
    except (KeyError, ValueError): return 0.0
agg = collections.OrderedDict()
launches = []
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0]
    name = name.replace("void ", "").split("<")[0].strip()   # template instantiations under one name
    a = agg.setdefault(name, collections.Counter())
    t = f(r, "gpu__time_duration.sum")
    launches.append((name, r[col["Grid Size"]], t))
    a["n"] += 1; a["t"] += t
    a["rd"] += f(r, "dram__bytes_read.sum"); a["wr"] += f(r, "dram__bytes_write.sum")
    a["inst"] += f(r, "smsp__inst_executed.sum")
    for k, m in (("issue", "smsp__issue_active.avg.pct_of_peak_sustained_active"), ("alu", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
                 ("fma", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"), ("lsu", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
                 ("occ", "sm__warps_active.avg.pct_of_peak_sustained_active")):
        a[k] += f(r, m) * t
unit_rd = rows[1][col["dram__bytes_read.sum"]]
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(unit_rd, 1e6)
tot = sum(a["t"] for a in agg.values())
print("| kernel | launches | time us | share | DRAM rd + wr MB | warp instr M | issue % | ALU % | FMA % | LSU % | warps active % |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
traffic = {}
for name, a in agg.items():
    t = a["t"]
    print("| %s | %d | %.0f | %.1f %% | %.1f + %.1f | %.1f | %.0f | %.0f | %.0f | %.0f | %.0f |" % (
        name, a["n"], t, 100 * t / tot, a["rd"] * scale / 1e6, a["wr"] * scale / 1e6, a["inst"] / 1e6, a["issue"] / t, a["alu"] / t, a["fma"] / t,
        a["lsu"] / t, a["occ"] / t))
    traffic[name] = {"dram_bytes_per_image": (a["rd"] + a["wr"]) * scale / imgs, "issue_slots_busy_pct": a["issue"] / t,
                     "alu_pipe_pct": a["alu"] / t, "us_per_image": t / imgs}
if len(sys.argv) > 3:
    traffic["_comment"] = "per-kernel sums from %s (%g images per launch): dram__bytes_read.sum + dram__bytes_write.sum per image, time-weighted issue / ALU utilisation" % (sys.argv[1].split("/")[-1], imgs)
    json.dump(traffic, open(sys.argv[3], "w"), indent=1)
if len(sys.argv) > 4:
    with open(sys.argv[4], "w") as fo:
        fo.write("kernel,grid,gpu__time_duration_us\n")
        for n, g, t in launches: fo.write('%s,"%s",%.3f\n' % (n, g, t))
