"""Summarise an `ncu --page source --csv` dump: instruction mix by opcode (weighted by executions) and the top stall sites.
Usage: python tools/ncu_source_summary.py <source.csv> [kernel-substring] [top-n]"""
import csv, sys, collections
path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
# split per kernel
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], hdr=None, rows=[])
        kernels.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r:
        cur["rows"].append(r)
for k in kernels:
    if want not in k["name"]:
        continue
    h = {n: i for i, n in enumerate(k["hdr"])}
    ex, smp, src = h["Instructions Executed"], h["# Samples"], h["Source"]
    tot_ex = sum(int(r[ex] or 0) for r in k["rows"])
    tot_s = sum(int(r[smp] or 0) for r in k["rows"])
    print(f"=== {k['name'][:90]}  instr={len(k['rows'])} executed={tot_ex} samples={tot_s}")
    mix = collections.Counter()
    smix = collections.Counter()
    for r in k["rows"]:
        toks = r[src].split()
        op = toks[0] if not toks[0].startswith("@") else toks[1]
        op = op.split(".")[0] if not op.startswith(("LDTM", "STTM", "UTC", "MUFU", "SYNCS")) else op
        mix[op] += int(r[ex] or 0)
        smix[op] += int(r[smp] or 0)
    print("opcode mix (share of executed warp-instructions | share of stall samples):")
    for op, c in mix.most_common(28):
        print(f"  {op:18s} {100*c/tot_ex:6.2f}%  | {100*smix[op]/max(tot_s,1):6.2f}%")
    stall_cols = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
    tots = {n: sum(float(r[h[n]] or 0) for r in k["rows"]) for n in stall_cols}
    s = sum(tots.values())
    print("stall reasons:", ", ".join(f"{n[6:]} {100*v/s:.1f}%" for n, v in sorted(tots.items(), key=lambda kv: -kv[1])[:8]))
    print(f"top {topn} sites by samples:")
    order = sorted(range(len(k["rows"])), key=lambda i: -int(k["rows"][i][smp] or 0))[:topn]
    for i in sorted(order):
        r = k["rows"][i]
        top = sorted(((float(r[h[n]] or 0), n[6:]) for n in stall_cols), reverse=True)[:2]
        print(f"  [{i:4d}] {100*int(r[smp] or 0)/max(tot_s,1):5.2f}%  ex={r[ex]:>9s}  {r[src].strip()[:70]:70s} {top[0][1]}:{top[0][0]:.0f} {top[1][1]}:{top[1][0]:.0f}")
