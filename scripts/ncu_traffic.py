"""Turn an `ncu --set full` capture into profiles/r02_ncu_traffic.json + a markdown table (run here, no GPU needed):
    ncu -i gpurun_out/<x>.ncu-rep --page raw --csv > /tmp/raw.csv ; python scripts/ncu_traffic.py /tmp/raw.csv <ctx> <out.md>
bench.py reads the JSON for `roofline.traffic` (DRAM bytes per launch of the dominant kernel)."""
import csv, json, os, sys
raw, ctx, out_md = sys.argv[1], float(sys.argv[2]), sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr = rows[0]
col = {n: i for i, n in enumerate(hdr)}
want = {"dram__bytes_read.sum": "rd", "dram__bytes_write.sum": "wr", "gpu__time_duration.sum": "dur",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_pct", "launch__registers_per_thread": "regs",
        "lts__t_sector_hit_rate.pct": "l2_hit", "launch__grid_size": "grid"}
units = rows[1]
def num(v):
    try: return float(v.replace(",", ""))
    except Exception: return None
def to_bytes(v, u):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return v * m.get(u, 1)
def to_us(v, u):
    m = {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}
    return v * m.get(u, 1)
recs = {}
lines = ["| kernel | grid | time us | DRAM read MB | DRAM write MB | DRAM % | L2 hit % | tensor pipe % | warps active % | regs |", "|---|---|---|---|---|---|---|---|---|---|"]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    name = r[col["Kernel Name"]]
    g = {k2: num(r[col[k]]) for k, k2 in want.items() if k in col}
    rd = to_bytes(g.get("rd") or 0, units[col["dram__bytes_read.sum"]])
    wr = to_bytes(g.get("wr") or 0, units[col["dram__bytes_write.sum"]])
    dur = to_us(g.get("dur") or 0, units[col["gpu__time_duration.sum"]])
    short = name.split("(")[0].replace("void ", "").replace("vcb::", "")
    lines.append(f"| `{short[:60]}` | {int(g.get('grid') or 0)} | {dur:.1f} | {rd/1e6:.1f} | {wr/1e6:.2f} | {g.get('dram_pct') or 0:.1f} | {g.get('l2_hit') or 0:.1f} | "
                 f"{g.get('tensor_pct') or 0:.1f} | {g.get('warps_pct') or 0:.1f} | {int(g.get('regs') or 0)} |")
    key = short.split("<")[0]
    recs.setdefault(key, []).append(dict(dram_bytes_per_launch=rd + wr, dram_read=rd, dram_write=wr, time_us=dur, ctx=ctx))
open(out_md, "w").write("\n".join(lines) + "\n")
js = {}
for k, v in recs.items():
    js[k] = dict(dram_bytes_per_launch=sum(x["dram_bytes_per_launch"] for x in v) / len(v), launches_captured=len(v), ctx=ctx,
                 time_us_under_ncu=sum(x["time_us"] for x in v) / len(v),
                 source=f"ncu --set full capture of this build ({os.path.basename(raw)}), dram__bytes_read.sum + dram__bytes_write.sum per launch")
json.dump(js, open(os.path.join(os.path.dirname(out_md) or ".", "r02_ncu_traffic.json"), "w"), indent=1)
print("\n".join(lines))
