"""Turns the ncu CSVs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1b.csv   > profiles/r01_launches.md
  python profiles/summarize.py raw gpurun_out/kernels_r1b_raw.csv     > profiles/r01_kernels.md
(raw csv = `ncu -i <rep> --page raw --csv`)
"""
import collections
import csv
import re
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi + 2:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"<.*", "", re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", ""))
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        v *= {"us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(unit, 1.0)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms (ncu, serialised, cold cache) | share |")
    print("|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.2f} % |")
    print(f"| total | {sum(v[0] for v in agg.values())} | {tot:.3f} | |")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"),
            ("launch__grid_size", "grid"), ("launch__waves_per_multiprocessor", "waves/SM"),
            ("sm__cycles_active.avg", "SM cycles active (avg)"), ("sm__cycles_elapsed.max", "cycles elapsed")]
    print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        vals = []
        for c, _ in cols:
            if c in idx:
                try:
                    vals.append(f"{float(r[idx[c]].replace(',', '')):.4g} {units[idx[c]]}")
                except ValueError:
                    vals.append(r[idx[c]])
            else:
                vals.append("-")
        print(f"| `{name[:60]}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
