"""Turns the scratch captures under gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarise_profiles.py launches gpurun_out/launches.csv profiles/r01_launches.md
    python tools/summarise_profiles.py raw gpurun_out/prof_X.ncu-rep profiles/r01_X_ncu.md
"""
import csv, io, re, subprocess, sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "lts__t_sectors_srcunit_tex.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__cycles_active.avg", "sm__cycles_active.avg",
]


def launches(src, dst):
    rows = [l for l in open(src) if l.startswith('"')]
    rd = csv.DictReader(io.StringIO("".join(rows)))
    agg = OrderedDict()
    for r in rd:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1]
        name = re.sub(r"<.*", "", name) if name.startswith("void") else name
        a = agg.setdefault(name, {"n": 0, "ns": 0.0, "grid": r["Grid Size"], "block": r["Block Size"], "full": r["Kernel Name"]})
        a["n"] += 1
        a["ns"] += float(r["Metric Value"].replace(",", ""))
    ours = {k: v for k, v in agg.items() if "sdfr::" in v["full"]}
    tot = sum(v["ns"] for v in ours.values())
    out = ["| kernel | launches | grid | block | total us | us / launch | share of our kernels |", "|---|---|---|---|---|---|---|"]
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1]["ns"]):
        out.append(f"| `{k}` | {v['n']} | {v['grid']} | {v['block']} | {v['ns']/1e3:,.1f} | {v['ns']/1e3/v['n']:,.1f} | {100*v['ns']/tot:.1f} % |")
    out.append("")
    out.append(f"Sum over our kernels: {tot/1e3:,.0f} us. Other kernels in the capture: " +
               ", ".join(f"`{k}` x{v['n']} ({v['ns']/1e3:,.0f} us)" for k, v in agg.items() if k not in ours) + ".")
    open(dst, "a").write("\n".join(out) + "\n")


def raw(src, dst):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = ["| metric | value | unit |", "|---|---|---|"]
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            out.append(f"| `{m}` | {vals[i]} | {units[i]} |")
    open(dst, "a").write("\n".join(out) + "\n")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
