"""Turn ncu exports into the small, committed summaries under profiles/:
  launches CSV (gpu__time_duration per launch)      -> per-kernel share table (markdown)
  --set full report of the tcgen05 kernel            -> per-launch table + tc_gemm_traffic.json
Usage: python tools/ncu_summary.py <launches.csv> <full.ncu-rep> <tag>
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launches(path, tag):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict(); tot = 0.0
    for r in rows:
        v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)      # -> us
        name = re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1]
        agg.setdefault(name, [0.0, 0]); agg[name][0] += v; agg[name][1] += 1; tot += v
    out = [f"# ncu launch list ({tag}): one 8-frame 512x512 chunk, frame -> code, bf16 operands",
           "# `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares)",
           "", "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, (v, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        out.append(f"| `{k}` | {n} | {v:.1f} | {100 * v / tot:.1f}% |")
    out.append(f"| **total** | {len(rows)} | {tot:.1f} | 100% |")
    return "\n".join(out) + "\n"


def full(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[0], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    want = [("dur_us", "gpu__time_duration.sum"), ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
            ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            ("regs", "launch__registers_per_thread"), ("smem_KB", "launch__shared_mem_per_block_dynamic")]
    units = rows[1]
    out = [f"# ncu --set full, tcgen05 implicit-GEMM kernel ({tag}): the 31 launches of one 8-frame 512x512 chunk", "",
           "| # | kernel | " + " | ".join(n for n, _ in want) + " |", "|---|---|" + "---:|" * len(want)]
    tot_bytes = 0.0
    for j, r in enumerate(data):
        name = re.search(r"tc_gemm_kernel<[^>]*>", r[col["Kernel Name"]])
        vals = []
        for n, m in want:
            v = float(r[col[m]].replace(",", "")); u = units[col[m]]
            if n == "dur_us":
                v = v * 1e3 if u == "ms" else (v / 1e3 if u == "ns" else v)
            if n.endswith("_MB"):
                v = v * {"Gbyte": 1e3, "Mbyte": 1, "Kbyte": 1e-3, "byte": 1e-6}[u]
                tot_bytes += v * 1e6
            if n == "smem_KB":
                v = v * {"Kbyte": 1, "byte": 1e-3, "Mbyte": 1e3}[u.split("/")[0]]
            vals.append(f"{v:.1f}")
        out.append(f"| {j} | `{name.group(0) if name else '?'}` | " + " | ".join(vals) + " |")
    traffic = dict(dram_bytes_per_launch=tot_bytes / max(len(data), 1), launches=len(data), source=os.path.basename(rep),
                   note="average of dram__bytes_read.sum + dram__bytes_write.sum over the tcgen05 launches of one 8-frame chunk")
    return "\n".join(out) + "\n", traffic


if __name__ == "__main__":
    lcsv, rep, tag = sys.argv[1:4]
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_launch_shares.md"), "w").write(launches(lcsv, tag))
    md, traffic = full(rep, tag)
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_tc_gemm_full.md"), "w").write(md)
    json.dump(traffic, open(os.path.join(ROOT, "profiles", "tc_gemm_traffic.json"), "w"), indent=1)
    print(md[:3000]); print(traffic)
