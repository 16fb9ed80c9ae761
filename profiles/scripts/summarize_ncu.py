"""Turns ncu output brought back in gpurun_out/ into the small summaries kept under profiles/.

  python profiles/scripts/summarize_ncu.py full  <rep.ncu-rep> <out.csv>     # --set full capture: key metrics + stall reasons per launch
  python profiles/scripts/summarize_ncu.py list  <launches.csv> <out.json>   # launch list (time + DRAM bytes per kernel) -> traffic.json
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__warps_eligible.avg.per_cycle_active"]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    with open(out, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["kernel"] + [f"{k} [{units[idx[k]]}]" for k in KEYS if k in idx] + ["top stall reasons (warps stalled per issue)"])
        for d in data:
            st = sorted(((float(d[idx[h]].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                         for h in stalls if d[idx[h]] not in ("", "n/a")), reverse=True)[:7]
            w.writerow([d[idx["Kernel Name"]].split("(")[0]] + [d[idx[k]] for k in KEYS if k in idx] + ["; ".join(f"{n} {v:.2f}" for v, n in st)])


def launch_list(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.defaultdict(dict)
    for r in rows[1:]:
        per[r[ii]]["k"] = r[ki]
        per[r[ii]][r[mi]] = float(r[vi].replace(",", ""))
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for v in per.values():
        name = v["k"].split("(")[0].replace("void ", "").replace("sllb::", "")
        a = agg[name]
        a[0] += 1; a[1] += v.get("gpu__time_duration.sum", 0.0); a[2] += v.get("dram__bytes_read.sum", 0.0); a[3] += v.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    res = {"source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on the bench command ({path}); "
                     "per-launch averages; times are cold-cache and serialised (shares, not bench values)", "kernels": {}}
    for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        res["kernels"][name] = {"launches": a[0], "us_per_launch": a[1] / a[0] / 1e3, "share_of_kernel_time": a[1] / tot,
                                "dram_bytes_per_launch": (a[2] + a[3]) / a[0], "dram_read_bytes_per_launch": a[2] / a[0],
                                "dram_write_bytes_per_launch": a[3] / a[0]}
    json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    {"full": full, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
