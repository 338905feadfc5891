"""Turn the ncu artefacts brought back in gpurun_out/ into the committed summaries under profiles/.
usage: python tools/make_profiles.py <round tag, e.g. r1>"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go = os.path.join(root, "gpurun_out")
pr = os.path.join(root, "profiles")
os.makedirs(pr, exist_ok=True)

# ---- launch list
src = os.path.join(go, f"launches_{tag}.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(pr, f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    start = rows.index(hdr) + 1
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[start:]:
        name = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(pr, f"{tag}_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 2 --warmup 1\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        for k, (c, t) in agg.items():
            f.write(f"{k[:64]:64s} launches={c:3d} total_ms={t:10.3f} avg_ms={t / c:9.3f} share={t / tot * 100:5.1f}%\n")
        f.write(f"total_ms={tot:.3f}\n")

# ---- full capture of the sort kernel
rep = os.path.join(go, f"bwt_{tag}_full.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units, vals = rows[0], rows[1], rows[2]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
    out = {}
    with open(os.path.join(pr, f"{tag}_bwt_sort_ncu.csv"), "w") as f:
        f.write("metric,unit,value\n")
        for n in want:
            if n in h:
                i = h.index(n)
                f.write(f"{n},{units[i]},{vals[i]}\n")
                out[n] = (units[i], vals[i])

    def to_bytes(u, v):
        v = float(v)
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)

    if "dram__bytes_read.sum" in out:
        traffic = to_bytes(*out["dram__bytes_read.sum"]) + to_bytes(*out["dram__bytes_write.sum"])
        tj = os.path.join(pr, "bwt_traffic.json")
        d = json.load(open(tj)) if os.path.exists(tj) else {}
        d["mixed-1GiB-L9"] = int(traffic)
        d["_source"] = f"profiles/{tag}_bwt_sort_ncu.csv (ncu --set full, one launch, bench workload)"
        json.dump(d, open(tj, "w"), indent=1)
    lines = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_lines.py"), rep, "bwt_sort_kernel",
                            os.path.join(root, "banzai_b200", "csrc", "bwt_sort.cu"), "40"],
                           capture_output=True, text=True).stdout
    open(os.path.join(pr, f"{tag}_bwt_sort_stalls_by_line.txt"), "w").write(
        "# warp-stall samples per source line (ncu --set full --import-source on, joined with nvdisasm -g)\n" + lines)

bj = os.path.join(go, f"bench_{tag}.json")
if os.path.exists(bj):
    shutil.copy(bj, os.path.join(pr, f"{tag}_bench.json"))
print(os.listdir(pr))
