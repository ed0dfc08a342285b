"""Aggregate an `ncu --page source --csv` SASS listing by CUDA source line (diagnostic tool).

    nvdisasm -g -c kernel.cubin > k.sass        # SASS with '//## File "...", line N' markers (needs -lineinfo)
    ncu -i rep.ncu-rep --page source --csv --launch-skip S --launch-count 1 > k.csv
    python tools/ncu_by_line.py k.sass k.csv [top]

The two listings are joined by instruction offset (ncu prints absolute addresses: the first row is offset 0).
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    sass, rep = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    line_of = {}
    cur = None
    for ln in open(sass):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(rep)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = None
    inst, samp = defaultdict(float), defaultdict(float)
    for r in rows[hi + 1:]:
        if r and r[0] == "Address":
            break  # a second table (another view of the same kernel) follows
        if len(r) <= ii or not r[ia]:
            continue
        addr = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        if base is None:
            base = addr
        key = line_of.get(addr - base)
        inst[key] += float(r[ii] or 0)
        samp[key] += float(r[isamp] or 0)
    ti, ts = sum(inst.values()), sum(samp.values())
    print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
    for key, v in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{str(key):38s} inst {100 * v / ti:5.1f}%  samples {100 * samp[key] / ts:5.1f}%")


if __name__ == "__main__":
    main()
