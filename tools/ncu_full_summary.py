"""Summarise an `ncu --set full` report (exported with `ncu -i X.ncu-rep --page raw --csv`) into a markdown table.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_full_summary.py /tmp/raw.csv "title" "command" > profiles/rNN_ncu_full.md
"""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "time us"),
    ("dram__bytes_read.sum", "dram rd MB"),
    ("dram__bytes_write.sum", "dram wr MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps act %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem KB"),
]


def main():
    path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    idx = [(hdr.index(m), label, units[hdr.index(m)]) for m, label in COLS if m in hdr]
    print(f"# {title}\n\nCommand: `{cmd}`\n(per-launch values under the profiler: serialised, cold caches; use for ratios and limiter analysis,"
          " never as a benchmark)\n")
    print("| kernel | " + " | ".join(l for _, l, _ in idx) + " |")
    print("|---|" + "---|" * len(idx))
    for r in data:
        name = r[ki].replace("advb::<unnamed>::", "").replace("(int)", "").replace("(bool)", "").split("(")[0]
        name = name.replace("void ", "")
        vals = []
        for i, label, unit in idx:
            v = r[i].replace(",", "")
            try:
                f = float(v)
                if unit == "byte":
                    f /= 1e6
                elif unit == "Kbyte":
                    f /= 1e3
                elif unit == "ns":
                    f /= 1e3
                elif unit == "ms":
                    f *= 1e3
                elif unit == "Gbyte":
                    f *= 1e3
                elif unit == "Kbyte/block":
                    pass
                vals.append(f"{f:.1f}" if f < 1e5 else f"{f:.0f}")
            except ValueError:
                vals.append(v)
        print(f"| `{name}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
