#!/usr/bin/env python
"""Per-launch table of an `ncu --set full ... --page raw --csv` dump of one engine plan (UNet step / VAE decode):
python tools/ncu_step_table.py raw.csv launches_per_plan [first_index_of_a_plan] > table.md
The capture may start in the middle of a plan: rows [first, first + n) are used, wrapping around n rows earlier if needed."""
import csv
import json
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    n = int(sys.argv[2])
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {k: i for i, k in enumerate(hdr)}

    def val(r, k):
        try:
            v = float(r[idx[k]].replace(",", ""))
        except (KeyError, ValueError):
            return 0.0
        u = units[idx[k]]
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "byte": 1e-6}.get(u, 1.0)
        return v * scale

    sel = []
    for j in range(n):
        i = first + j
        if i >= len(data):
            i -= n
        sel.append(data[i])
    T, TP = "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    print("| # | kernel | grid | regs | us | tensor pipe % (active) | DRAM read MB | DRAM write MB | DRAM % of peak | L2 hit % | warps active % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|")
    fam = {}
    for j, r in enumerate(sel):
        name = r[idx["Kernel Name"]].replace("void ", "").split("(")[0][:44]
        t = val(r, T)
        f = fam.setdefault(name, dict(n=0, us=0.0, tp=0.0, rd=0.0, wr=0.0))
        f["n"] += 1; f["us"] += t; f["tp"] += t * val(r, TP); f["rd"] += val(r, "dram__bytes_read.sum"); f["wr"] += val(r, "dram__bytes_write.sum")
        print(f"| {j} | `{name}` | {int(val(r, 'launch__grid_size'))} | {int(val(r, 'launch__registers_per_thread'))} | {t:.1f} | "
              f"{val(r, TP):.1f} | {val(r, 'dram__bytes_read.sum'):.1f} | {val(r, 'dram__bytes_write.sum'):.1f} | "
              f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {val(r, 'lts__t_sector_hit_rate.pct'):.1f} | "
              f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} |")
    tot = sum(f["us"] for f in fam.values())
    print("\n| kernel family | launches | total us | share | time-weighted tensor pipe % | DRAM MB per launch (read + write) | achieved DRAM GB/s |")
    print("|---|---|---|---|---|---|---|")
    for name, f in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        print(f"| `{name}` | {f['n']} | {f['us']:.1f} | {100 * f['us'] / tot:.1f} % | {f['tp'] / max(f['us'], 1e-9):.1f} | "
              f"{(f['rd'] + f['wr']) / f['n']:.1f} | {(f['rd'] + f['wr']) / max(f['us'], 1e-9) * 1e3:.0f} |")
    conv = {k: v for k, v in fam.items() if "conv_tc" in k}
    if conv:
        nl = sum(v["n"] for v in conv.values())
        b = sum(v["rd"] + v["wr"] for v in conv.values()) * 1e6 / nl
        print("\nroofline_traffic: " + json.dumps({"dram_bytes_per_launch": b, "launches_captured": nl}))


if __name__ == "__main__":
    main()
