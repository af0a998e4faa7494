"""Summarise ncu artefacts into small text files for profiles/.
  python scripts/ncu_summary.py launches <launches.csv>        -> per-kernel share of the captured window
  python scripts/ncu_summary.py full <file.ncu-rep>            -> key counters of each captured launch
  python scripts/ncu_summary.py traffic <fwd.ncu-rep> <tag>    -> profiles/<tag>_traffic.json (DRAM bytes per NT GEMM launch)
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki]))
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# %s : %d launches, %.1f us total (ncu per-launch times are cold-cache and serialised: compare SHARES)" % (path, sum(a[0] for a in agg.values()), tot))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s n=%4d %10.1f us %6.2f%%  avg %8.1f us" % (k[:72], a[0], a[1], 100 * a[1] / tot, a[1] / a[0]))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print("## " + d.get("Kernel Name", "?")[:120])
        for k in KEYS:
            if k in d:
                print("  %-78s %14s %s" % (k, d[k], units[hdr.index(k)]))


def traffic(path, tag):
    """DRAM bytes per launch of the dominant kernel family (pair NT GEMM) from a --set full capture -> profiles/<tag>_traffic.json
    (read by bench.py for roofline.traffic).  The forward capture holds block 0's QKV, proj, fc1, fc2 launches in that order."""
    import json
    import os
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    names = ["QKV N=1536 K=512", "proj N=512 K=512 (+fp32 residual)", "fc1 N=2048 K=512", "fc2 N=512 K=2048 (+fp32 residual)"]
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    ls = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if "gemm_nt3" not in d.get("Kernel Name", ""):
            continue
        rd = float(d["dram__bytes_read.sum"].replace(",", "")) * mult[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(d["dram__bytes_write.sum"].replace(",", "")) * mult[units[hdr.index("dram__bytes_write.sum")]]
        ls.append({"what": names[len(ls)] if len(ls) < len(names) else "?", "dram_read_mb": rd / 1e6, "dram_write_mb": wr / 1e6,
                   "duration_us": float(d["gpu__time_duration.sum"].replace(",", ""))})
    ls = ls[:4]
    res = {"kernel": "gemm_nt3_kernel<256, 5>",
           "source": "profiles/%s_top_kernels_full.txt (ncu --set full, cold cache, bench.py --no-graph --quick --steps 1)" % tag,
           "launches": ls,
           "avg_dram_bytes_per_launch": int(sum((x["dram_read_mb"] + x["dram_write_mb"]) * 1e6 for x in ls) / max(1, len(ls))),
           "note": "the first four NT GEMM launches of the step (block 0 forward); writes mostly stay in the 126 MB L2 for the duration "
                   "of one launch, so dram__bytes_write under-counts the output traffic"}
    dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "%s_traffic.json" % tag)
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
