#!/usr/bin/env python
"""Turn the ncu artefacts fetched from the GPU box (gpurun_out/) into the tracked summaries under profiles/."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)
short = lambda n: n.split("(")[0].replace("<unnamed>::", "").replace("void ", "").strip()[:60]
# 1. launch list
rows = [r for r in csv.reader(open(os.path.join(G, f"{R}_launches.csv"))) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
with open(os.path.join(P, f"{R}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --quick --steps 4 --warmup 3\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES of the step, not absolutes\n")
    f.write("launch,kernel,duration_ns\n")
    for i, r in enumerate(rows[1:]):
        f.write(f"{i},{short(r[ki])},{r[vi]}\n")
tot = {}
for r in rows[1:]:
    tot.setdefault(short(r[ki]), []).append(float(r[vi].replace(",", "")))
share = {k: sum(v) for k, v in tot.items()}
ours = {k: v for k, v in share.items() if "kernel" in k}
s = sum(ours.values())
# 2. full capture -> raw csv of chosen metrics
rep = os.path.join(G, f"{R}_kernels.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h, u = rr[0], rr[1]
idx = {n: i for i, n in enumerate(h)}
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__inst_executed.max", "smsp__inst_executed.avg",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_bytes.sum", "gpc__cycles_elapsed.avg.per_second"]
keep += [n for n in h if "issue_stalled" in n and "per_issue_active" in n and "not_issued" not in n]
keep = [k for k in keep if k in idx]
with open(os.path.join(P, f"{R}_kernels_raw.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(keep); w.writerow([u[idx[k]] for k in keep])
    for r in rr[2:]:
        w.writerow([r[idx[k]] for k in keep])
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return None
def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * m.get(unit, 1)
dram = {}
for r in rr[2:]:
    name = short(r[idx["Kernel Name"]]).replace("_kernel", "").replace("blend_backward_pipe", "blend_backward")
    b = to_bytes(num(r[idx["dram__bytes_read.sum"]]), u[idx["dram__bytes_read.sum"]]) + to_bytes(num(r[idx["dram__bytes_write.sum"]]), u[idx["dram__bytes_write.sum"]])
    dram[name] = b
summary = {"round": R, "command": "python bench.py --quick --steps 4 --warmup 3 (config 2: 512x512, 100k Gaussians)",
           "share_of_step_pct": {k: round(100 * v / s, 1) for k, v in sorted(ours.items(), key=lambda kv: -kv[1])},
           "dram_bytes_per_launch": dram}
json.dump(summary, open(os.path.join(P, f"{R}_summary.json"), "w"), indent=1)
print(json.dumps(summary, indent=1))
