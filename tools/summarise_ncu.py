"""Turns gpurun_out/launches.csv + gpurun_out/prof_r1.ncu-rep into the text summaries committed under profiles/."""
import collections, csv, json, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 5]
hdr, data = None, []
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr:
        data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    k = re.sub(r"\(.*", "", d["Kernel Name"])[:80]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(d["Metric Value"].replace(",", ""))
out = ["# ncu launch list (gpu__time_duration.sum, --clock-control none): python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e",
       "# per-launch times are cold-cache and serialised by ncu: compare SHARES.  share_frame = share among the three render kernels."]
tot = sum(a[1] for a in agg.values())
steady = {k: a for k, a in agg.items() if re.search(r"k_primary_fast_simple<0>|k_shade<0, 0|k_write", k)}
ts = sum(a[1] for a in steady.values())
for k, a in agg.items():
    out.append(f"{k:82s} launches={a[0]:4d} total_us={a[1]/1e3:10.1f} avg_us={a[1]/a[0]/1e3:9.2f} share_all={a[1]/tot:.3f}" +
               (f" share_frame={a[1]/ts:.3f}" if k in steady else ""))
open(f"profiles/{tag}_ncu_launch_list_summary.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out[-8:]))
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_r1.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
o = ["# ncu --set full --clock-control none, one launch of each render kernel (config 3, N=1); from gpurun_out/prof_r1.ncu-rep"]
traffic = {}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for r in rows[2:5]:
    name = r[idx["Kernel Name"]]
    o.append("---- " + name[:120])
    for w in want:
        if w in idx:
            o.append(f"  {w:72s} {r[idx[w]]} {units[idx[w]]}")
    for h in hdr:
        if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct"):
            try:
                if float(r[idx[h]]) > 2:
                    o.append("   stall " + h.replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", "") + " " + r[idx[h]])
            except ValueError:
                pass
    tb = sum(float(r[idx[m]]) * mult[units[idx[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    traffic["primary" if "k_primary" in name else ("shade" if "k_shade" in name else "write")] = tb
open(f"profiles/{tag}_ncu_full_summary.txt", "w").write("\n".join(o) + "\n")
json.dump(traffic, open("profiles/traffic.json", "w"))
print("\n".join(o))
print(traffic)
