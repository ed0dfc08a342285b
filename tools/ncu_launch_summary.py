"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (per-kernel share).

    python tools/ncu_launch_summary.py gpurun_out/launches.csv "title" "command" > profiles/rNN_launches.md
"""
import collections
import csv
import sys


def main():
    path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        name = row["Kernel Name"].replace("unnamed>::", "").replace("void ", "")
        name = name.split("(const")[0].split("(TcArgs")[0].split("(float")[0]
        a = agg.setdefault(name, [0, 0.0, row["Grid Size"], row["Block Size"]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {title}\n\nCommand: `{cmd}`\n(cold-cache serialised per-launch times: compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | avg us | share | grid | block |\n|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.1f} | {100 * a[1] / tot:.1f}% | {a[2]} | {a[3]} |")
    print(f"\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")


if __name__ == "__main__":
    main()
