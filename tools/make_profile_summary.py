"""Regenerate profiles/<tag>_ncu_summary.md and profiles/gemm_traffic.json from the ncu exports a tools/gpu_round.sh run
left in gpurun_out/ (launch list + `--set full` raw pages).  Usage: python tools/make_profile_summary.py [tag]"""
import collections
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second"]


def raw(kind):
    path = os.path.join(SRC, f"{TAG}_prof_{kind}_raw.csv")
    if not os.path.isfile(path):
        return None, None, []
    shutil.copy(path, os.path.join(DST, f"{TAG}_ncu_{kind}_raw.csv"))
    rows = list(csv.reader(open(path)))
    return rows[0], rows[1], rows[2:]


def table(hdr, units, row):
    h = {n: i for i, n in enumerate(hdr)}
    out = ["| metric | value | unit |", "|---|---|---|", f"| Kernel Name | `{row[h['Kernel Name']][:110]}` | |"]
    for k in KEYS:
        if k in h:
            out.append(f"| {k} | {row[h[k]]} | {units[h[k]]} |")
    return "\n".join(out)


def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


md = [f"# {TAG}: ncu evidence (`--set full --clock-control none`, `tools/prof_kernels.py`, BASELINE config C2 shapes)", "",
      "Read with `ncu -i ... --page raw --csv`; the raw pages are committed next to this file.  Times under ncu are cold-cache and",
      "serialised: compare SHARES with bench.py, not absolutes.", ""]

# ---- launch list
lpath = os.path.join(SRC, f"{TAG}_launches.csv")
if os.path.isfile(lpath):
    shutil.copy(lpath, os.path.join(DST, f"{TAG}_ncu_launches.csv"))
    rows = list(csv.reader(open(lpath)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = {n: i for i, n in enumerate(rows[hi])}
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        name = r[h["Kernel Name"]].split("(")[0].replace("void ", "")[:50]
        t = float(r[h["Metric Value"]]) / 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    md += [f"## Launch list (`ncu --metrics gpu__time_duration.sum`, `python bench.py --steps 1 --warmup 3`, {sum(a[0] for a in agg.values())} launches after the warm-up steps)", "",
           "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % |")
    md.append("")

# ---- GEMM
hdr, units, rows = raw("gemm")
names = ["QKV 123000x2304x768 bias", "proj 123000x768x768 +residual", "fc1 123000x3072x768 GELU", "fc2 123000x768x3072 +residual"]
if rows:
    h = {n: i for i, n in enumerate(hdr)}
    per = []
    for nm, r in zip(names, rows[:4]):
        md += [f"## GEMM {nm}", "", table(hdr, units, r), ""]
        rd = to_bytes(r[h["dram__bytes_read.sum"]], units[h["dram__bytes_read.sum"]])
        wr = to_bytes(r[h["dram__bytes_write.sum"]], units[h["dram__bytes_write.sum"]])
        per.append(dict(kernel=nm, dram_read_bytes=rd, dram_write_bytes=wr, dram_bytes=rd + wr,
                        duration_us=float(r[h["gpu__time_duration.sum"]])))
    json.dump(dict(source=f"profiles/{TAG}_ncu_gemm_raw.csv (ncu --set full --clock-control none, tools/prof_kernels.py gemm, 3rd launch of each shape)",
                   dram_bytes_per_launch=sum(p["dram_bytes"] for p in per) / len(per),
                   note="mean over the four per-layer GEMM shapes of config C2 (each launched 12x per step); dram__bytes_read.sum + dram__bytes_write.sum",
                   per_shape=per), open(os.path.join(DST, "gemm_traffic.json"), "w"), indent=1)
for kind, title in (("attn", "Flash attention B=120 N=1025 heads=12"), ("ln", "LayerNorm 123000x768 bf16"), ("gather", "Mask gather (C2 ellipsoid mask, then a dense mask)"),
                    ("sam", "MedSAM attention, 4 images of 64x64 tokens, 12 heads: global block (fused rel-pos flash kernel), then a windowed block (tcgen05 14x14 windows)")):
    hdr, units, rows = raw(kind)
    for r in rows:
        md += [f"## {title}", "", table(hdr, units, r), ""]
extra = os.path.join(DST, f"{TAG}_reading.md")
if os.path.isfile(extra):
    md += [open(extra).read()]
open(os.path.join(DST, f"{TAG}_ncu_summary.md"), "w").write("\n".join(md) + "\n")
print("wrote", os.path.join(DST, f"{TAG}_ncu_summary.md"))
