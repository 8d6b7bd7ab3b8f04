"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/perf/summarize_ncu.py launches gpurun_out/launches_r1d.csv  > profiles/r1_launches.txt
    python tools/perf/summarize_ncu.py kernels  gpurun_out/prof_r1d.ncu-rep  > profiles/r1_top_kernels.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    kn, mv = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[kn], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none launch list (cold-cache, serialised: compare SHARES)")
    print("# source: %s ; total kernel time %.3f ms over %d launches" % (path, tot / 1e6, sum(v[0] for v in agg.values())))
    print("%-72s %7s %12s %8s" % ("kernel", "calls", "ms", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s %7d %12.3f %7.2f%%" % (k[:72], v[0], v[1] / 1e6, 100 * v[1] / tot))


def kernels(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none capture, selected raw metrics per profiled launch; source: %s" % path)
    for r in rows[2:]:
        d = dict(zip(h, r))
        print("\n== %s  (ID %s)" % (d.get("Kernel Name", "?")[:100], d.get("ID", "?")))
        for k in KEYS:
            if k in d:
                print("  %-88s %s %s" % (k, d[k], units[h.index(k)]))
        try:
            rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
            ns = float(d["gpu__time_duration.sum"])
            print("  %-88s %.0f byte  (%.1f GB/s over this launch)" % ("traffic = dram__bytes_read.sum + dram__bytes_write.sum", rd + wr, (rd + wr) / ns))
        except (KeyError, ValueError):
            pass


def stalls(path):
    """Warp-stall shares and executed instructions of the one profiled kernel, from the SASS source page."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
    h, data = rows[hi], rows[hi + 1:]
    iex, ismp = h.index("Instructions Executed"), h.index("# Samples")
    tot = sum(int(r[ismp]) for r in data) or 1
    print("# SASS page of %s: %d instructions, %d warp-level instructions executed" % (path, len(data), sum(int(r[iex]) for r in data)))
    for name in [x for x in h if x.startswith("stall_") and "Not Issued" not in x]:
        share = 100.0 * sum(int(r[h.index(name)]) for r in data) / tot
        if share >= 1.0:
            print("  %-28s %5.1f %% of warp samples" % (name, share))


def traffic(spec):
    """traffic stark:kernel_regex=report.ncu-rep ... -> JSON {stark: {kernel: bytes per launch}} on stdout."""
    import json
    import re
    res = {}
    for item in spec:
        key, path = item.split("=")
        stark, pat = key.split(":")
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h = rows[0]
        for r in rows[2:]:
            d = dict(zip(h, r))
            m = re.search(pat, d.get("Kernel Name", ""))
            if m:
                name = m.group(0)
                res.setdefault(stark, {})[name] = int(float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"]))
    print(json.dumps(res, indent=1, sort_keys=True))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2:])
    else:
        {"launches": launches, "kernels": kernels, "stalls": stalls}[sys.argv[1]](sys.argv[2])
