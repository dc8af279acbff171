"""Copy the round's evidence from gpurun_out/ (scratch) into profiles/ (tracked): bench lines, parity curves, launch
lists + per-kernel step shares, the c5 update sweep, and the counters of the two `ncu --set full` captures.
    python tools/collect_profiles.py [round-tag, default r2]"""
import csv, io, json, os, shutil, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import launch_summary

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"


def last_json_line(path):
    if not os.path.exists(path):
        return None
    for l in reversed(open(path).read().strip().splitlines()):
        if l.startswith("{"):
            return json.loads(l)
    return None


def cp(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
        print("copied", src, "->", dst)


for src, dst in (("bench_full.log", f"{tag}_bench_n1.json"), ("bench_ref.log", f"{tag}_bench_reference_arm.json"),
                 ("bench_attn.log", f"{tag}_bench_attn_extension_n1.json"), ("bench_c2.log", f"{tag}_bench_c2.json"),
                 ("bench_c4.log", f"{tag}_bench_c4_shard.json"), ("bench_n2.log", f"{tag}_bench_n2.json"),
                 ("bench_n4.log", f"{tag}_bench_n4.json"), ("bench_n8.log", f"{tag}_bench_n8.json"),
                 ("bench_n8_weak.log", f"{tag}_bench_n8_weak.json"), ("bench_n8_c4.log", f"{tag}_bench_n8_c4.json")):
    d = last_json_line(os.path.join(G, src))
    if d:
        json.dump(d, open(os.path.join(P, dst), "w"), indent=1)
        print("wrote", dst)
cp("parity_growth.json", f"{tag}_parity_growth.json")
cp("parity_report.jsonl", f"{tag}_parity_report.jsonl")
cp("update_sweep.json", f"{tag}_update_sweep_c5.json")
cp("launches_ref.csv", f"{tag}_launches_ncu_ref.csv")
cp("launches_attn.csv", f"{tag}_launches_ncu_attn.csv")
cp("conv_layers.log", f"{tag}_conv_layers.txt")
cp("attn_bench.log", f"{tag}_attn_core.txt")
for n in ("layers_f16_a.log", "layers_bf16_a.log", "layers_f16_b.log", "layers_bf16_b.log", "layers_f32_1.log", "layers_half_1.log",
          "layers_f32_2.log", "layers_half_2.log", "layers_f32_c8.log"):
    cp(n, f"{tag}_ab_{n.replace('.log', '.txt')}")
for src, dst in (("bench_b8.log", f"{tag}_bench_n1_batch8.json"),):
    d = last_json_line(os.path.join(G, src))
    if d:
        json.dump(d, open(os.path.join(P, dst), "w"), indent=1)
with open(os.path.join(P, f"{tag}_launch_shares.txt"), "w") as f:
    old = sys.stdout
    sys.stdout = f
    for n in ("launches_ref.csv", "launches_attn.csv"):
        if os.path.exists(os.path.join(G, n)):
            launch_summary.main(os.path.join(G, n))
    sys.stdout = old

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_active", "sm__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
traffic, md = {}, [f"# {tag}: counters of the `ncu --set full --clock-control none` captures (tools/gpu/round.sh)\n"]
for rep in (f"prof_conv_128_{tag}", f"prof_update_{tag}", f"prof_conv_64_{tag}", f"prof_attn_{tag}"):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units, d = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(h)}
    name = d[col["Kernel Name"]].replace("void ", "").split("(")[0]
    md.append(f"\n## {rep}: `{name}`\n\n| counter | value |\n|---|---|")
    vals = {}
    for w in WANT:
        if w in col:
            vals[w] = (d[col[w]], units[col[w]])
            md.append(f"| {w} | {d[col[w]]} {units[col[w]]} |")

    def byt(k):
        v, u = vals[k]
        return float(v.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
    traffic[name] = {"traffic_MB": (byt("dram__bytes_read.sum") + byt("dram__bytes_write.sum")) / 1e6,
                     "read_MB": byt("dram__bytes_read.sum") / 1e6, "write_MB": byt("dram__bytes_write.sum") / 1e6,
                     "duration_us_under_ncu": float(vals["gpu__time_duration.sum"][0].replace(",", "")), "capture": rep}
rot = os.path.join(G, "upd_traffic_rotating.csv")
if os.path.exists(rot):  # steady-state DRAM bytes per launch of the update step (write-back included)
    rows = list(csv.reader(open(rot)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    col = {n: i for i, n in enumerate(rows[hi])}
    tot, n = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0}, 0
    for r in rows[hi + 1:]:
        if len(r) == len(rows[hi]) and r[col["Metric Name"]] in tot:
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[r[col["Metric Unit"]]]
            tot[r[col["Metric Name"]]] += float(r[col["Metric Value"]].replace(",", "")) * mult
            n += 1
    if n:
        launches = n // 2
        for k in list(traffic):
            if k.startswith("superpose_update"):
                traffic[k].update({"traffic_MB": (tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"]) / launches / 1e6,
                                   "read_MB": tot["dram__bytes_read.sum"] / launches / 1e6,
                                   "write_MB": tot["dram__bytes_write.sum"] / launches / 1e6,
                                   "how": f"mean of {launches} launches deep in the rotating sequence, --cache-control none"})
if traffic:
    json.dump(traffic, open(os.path.join(P, f"{tag}_traffic.json"), "w"), indent=1)
    open(os.path.join(P, f"{tag}_ncu.md"), "w").write("\n".join(md) + "\n")
    print("wrote", f"{tag}_traffic.json", f"{tag}_ncu.md")
