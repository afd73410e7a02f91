#!/usr/bin/env python
"""Summarise an .ncu-rep (read offline with `ncu -i`) into a short markdown table: python tools/ncu_summary.py rep [out.md]"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("gpu__time_duration.sum", "time_ns"),
    ("sm__cycles_elapsed.avg", "sm_cycles"), ("sm__cycles_active.avg", "sm_active_cycles"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct_active"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor_hmma_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_insts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "mem_throughput_pct"),
    ("dram__bytes_read.sum", "dram_read_B"), ("dram__bytes_write.sum", "dram_write_B"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_bytes.sum", "l2_bytes"), ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall_long_sb"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall_barrier"),
    ("smsp__warp_issue_stalled_membar_per_warp_active.pct", "stall_membar"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {k: i for i, k in enumerate(hdr)}
    lines = ["| " + " | ".join(n for _, n in KEYS if _ in idx) + " |", "|" + "---|" * sum(1 for k, _ in KEYS if k in idx)]
    for r in rows[2:]:
        cells = []
        for k, _ in KEYS:
            if k in idx:
                v = r[idx[k]]
                if k == "Kernel Name":
                    v = v[:40]
                cells.append(f"{v} {units[idx[k]]}".strip())
        lines.append("| " + " | ".join(cells) + " |")
    extra = [h for h in hdr if "tensor" in h.lower() or "tmem" in h.lower() or "utc" in h.lower()]
    text = "\n".join(lines) + "\n\nmetrics mentioning tensor/tmem: " + ", ".join(extra[:40]) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
