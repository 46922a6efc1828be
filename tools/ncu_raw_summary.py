"""Summarise `ncu -i X.ncu-rep --page raw --csv` (exported on the GPU box; the .ncu-rep itself is too large to
bring back) into profiles/<tag>_ncu_full.md and profiles/tc_gemm_traffic.json.
Usage: python tools/ncu_raw_summary.py gpurun_out/<raw>.csv <tag>
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [("dur_us", "gpu__time_duration.sum"),
        ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("regs", "launch__registers_per_thread"), ("smem_KB", "launch__shared_mem_per_block_dynamic")]
SCALE = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}


def main(path, tag):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = [f"# ncu --set full --clock-control none ({tag}): every tcgen05 / GroupNorm / conv_in / softmax launch of one",
           "# 8-frame 512x512 chunk (tools/profile_step.py mixed 8: mixed operand mode, 16-bit residual stream).  Per-launch times are cold-cache and",
           "# serialised under the profiler: use the shares and the per-launch traffic, not the absolute times.", "",
           "| # | kernel | " + " | ".join(n for n, _ in WANT) + " |", "|---|---|" + "---:|" * len(WANT)]
    agg = {}
    for j, r in enumerate(data):
        full = r[col["Kernel Name"]]
        m = re.search(r"(\w+_kernel(<[^>]*>)?)", full)
        name = m.group(1) if m else full[:40]
        vals = {}
        for n, metric in WANT:
            if metric not in col or r[col[metric]] in ("", "n/a"):
                vals[n] = float("nan"); continue
            v = float(r[col[metric]].replace(",", "")); u = units[col[metric]]
            if n == "dur_us":
                v = v * 1e3 if u == "ms" else (v / 1e3 if u == "ns" else v)
            elif n.endswith("_MB"):
                v *= SCALE[u]
            elif n == "smem_KB":
                v *= {"Kbyte": 1.0, "byte": 1e-3, "Mbyte": 1e3}[u.split("/")[0]]
            vals[n] = v
        out.append(f"| {j} | `{name}` | " + " | ".join(f"{vals[n]:.1f}" for n, _ in WANT) + " |")
        cls = re.sub(r"<.*", "", name)
        if name.startswith("tc_gemm_kernel<128, 1, 0>"):
            cls = "tc_gemm_kernel<128,1,0> (conv_in mode)"       # uint8-fed first layer: write-bound, accounted as conv_in
        a = agg.setdefault(cls, dict(launches=0, us=0.0, bytes=0.0, tensor_w=0.0))
        a["launches"] += 1; a["us"] += vals["dur_us"]; a["bytes"] += (vals["dram_rd_MB"] + vals["dram_wr_MB"]) * 1e6
        if vals["tensor_pct"] == vals["tensor_pct"]:
            a["tensor_w"] += vals["tensor_pct"] * vals["dur_us"]
    tot = sum(a["us"] for a in agg.values())
    out += ["", "| kernel class | launches | total us | share | DRAM bytes / launch | duration-weighted tensor_pct |",
            "|---|---:|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        out.append(f"| `{k}` | {a['launches']} | {a['us']:.1f} | {100 * a['us'] / tot:.1f}% | "
                   f"{a['bytes'] / a['launches'] / 1e6:.1f} MB | {a['tensor_w'] / a['us']:.1f} |")
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full.md"), "w").write("\n".join(out) + "\n")
    t = agg.get("tc_gemm_kernel")
    if t:
        traffic = dict(dram_bytes_per_launch=t["bytes"] / t["launches"], launches=t["launches"], source=os.path.basename(path),
                       note="average of dram__bytes_read.sum + dram__bytes_write.sum over the tcgen05 launches of one "
                            "8-frame 512x512 chunk (ncu --set full)",
                       per_class={k: dict(dram_bytes_per_launch=a["bytes"] / a["launches"], launches=a["launches"])
                                  for k, a in agg.items()})
        json.dump(traffic, open(os.path.join(ROOT, "profiles", "tc_gemm_traffic.json"), "w"), indent=1)
        print(traffic)
    print("\n".join(out[-8:]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
