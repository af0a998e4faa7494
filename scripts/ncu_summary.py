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


def families(path, tag):
    """Launch list with several metrics per launch (scripts/r02_profile_final.sh) -> per kernel: launches, time, DRAM bytes, tensor-pipe
    activity, executed bf16 tensor math ops; writes profiles/<tag>_traffic.json for the NT GEMM family (read by bench.py)."""
    import json
    import os
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, ni, vi, ui, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki]))
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        u = r[ui]
        mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
        per.setdefault((r[ii], name), {})[r[ni]] = v * mult
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        a = agg.setdefault(name, collections.Counter())
        a["n"] += 1
        a["us"] += m.get("gpu__time_duration.sum", 0.0)
        a["dram"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a["ops"] += m.get("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32.sum", 0.0)
        t = m.get("gpu__time_duration.sum", 0.0)
        a["pipe_w"] += t * m.get("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["hmma_w"] += t * m.get("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 0.0)
        a["memt_w"] += t * m.get("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
    tot = sum(a["us"] for a in agg.values())
    print("# %s : %d launches, %.1f us total (ncu: cold cache, serialised launches -> compare SHARES, not absolutes)" % (path, sum(a["n"] for a in agg.values()), tot))
    print("# tensor = sm__pipe_tensor_cycles_active_realtime %% of peak (time-weighted), hmma = ..._subpipe_hmma_..., memT = sm__mem_tensor_cycles_active %%;")
    print("# TF/s = sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32 (executed tensor math ops, FMA = 2) / kernel time")
    print("%-62s %5s %10s %7s %9s %7s %7s %7s %8s" % ("kernel", "n", "us", "share", "MB/launch", "tensor", "hmma", "memT", "TF/s"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        us = max(a["us"], 1e-9)
        print("%-62s %5d %10.1f %6.2f%% %9.2f %6.1f%% %6.1f%% %6.1f%% %8.1f" % (k[:62], a["n"], a["us"], 100 * a["us"] / tot, a["dram"] / a["n"] / 1e6,
                                                                     a["pipe_w"] / us, a["hmma_w"] / us, a["memt_w"] / us, a["ops"] / us / 1e6))
    nt = [a for k, a in agg.items() if "gemm_nt" in k]
    if nt:
        n = sum(a["n"] for a in nt)
        res = {"family": "gemm_bf16_nt", "kernels": [k for k in agg if "gemm_nt" in k], "launches": int(n),
               "avg_dram_bytes_per_launch": int(sum(a["dram"] for a in nt) / n),
               "source": "profiles/%s_launches_metrics.txt (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum over every launch of one eager step of the "
                         "default workload; cold cache; writes that stay in the 126 MB L2 are not counted)" % tag}
        dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "%s_traffic.json" % tag)
        json.dump(res, open(dst, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "families":
        families(sys.argv[2], sys.argv[3])
        sys.exit(0)
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], sys.argv[3])
    else:
        {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
