"""Turns an `ncu --set full` report into (a) a text summary for profiles/ and (b) the per-kernel physical numbers of
profiles/traffic.json that bench.py puts beside its roofline (DRAM / L2 bytes per launch, issue-slot and L1-pipe utilisation).

  python tools/summarise_ncu.py gpurun_out/r02_prof_c3.ncu-rep c3 profiles/r02_ncu_c3_summary.txt ["header line"]
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum", "SM_B.TriageCompute.l1tex__t_sectors.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
MULT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short_name(full):
    m = re.search(r"(k_[A-Za-z0-9_]+)", full)
    return m.group(1) if m else full[:40]


def main():
    rep, config, out_path = sys.argv[1], sys.argv[2], sys.argv[3]
    header = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = ["# ncu --set full --clock-control none --import-source on; one launch per kernel; report: %s" % os.path.basename(rep)]
    if header:
        out.append("# " + header)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    if not isinstance(traffic.get(config), dict):
        traffic[config] = {}

    def val(r, m):
        v = float(r[idx[m]].replace(",", ""))
        return v * MULT.get(units[idx[m]], 1.0)
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        out.append("---- " + name[:130])
        for w in WANT:
            if w in idx:
                out.append("  %-78s %s %s" % (w, r[idx[w]], units[idx[w]]))
        stalls = []
        for h in hdr:
            m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
            if m:
                try:
                    stalls.append((float(r[idx[h]]), m.group(1)))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        out.append("  warps stalled per issue-active cycle: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]))
        traffic[config][short_name(name)] = {
            "dram_bytes": val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
            "l2_bytes": 32.0 * val(r, "lts__t_sectors.sum") if "lts__t_sectors.sum" in idx else None,                 # 32-byte sectors
            "l1_bytes": 32.0 * val(r, "SM_B.TriageCompute.l1tex__t_sectors.sum") if "SM_B.TriageCompute.l1tex__t_sectors.sum" in idx else None,
            "l2_to_l1_read_bytes": val(r, "l1tex__m_xbar2l1tex_read_bytes.sum") if "l1tex__m_xbar2l1tex_read_bytes.sum" in idx else None,
            "duration_us_under_ncu": val(r, "gpu__time_duration.sum") / (1.0 if units[idx["gpu__time_duration.sum"]] == "us" else 1e3) if units[idx["gpu__time_duration.sum"]] in ("us", "ns") else None,
            "inst_executed": val(r, "smsp__inst_executed.sum"),
            "threads_per_inst": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "l1tex_throughput_pct": val(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers": val(r, "launch__registers_per_thread"),
            "source": os.path.relpath(out_path, ROOT),
        }
    open(out_path, "w").write("\n".join(out) + "\n")
    json.dump(traffic, open(tpath, "w"), indent=1, sort_keys=True)
    print("\n".join(out))


if __name__ == "__main__":
    main()
