#!/usr/bin/env python
"""Condense ncu output into the small, tracked summaries kept under profiles/.

    python tools/ncu_summary.py launches <launches.csv> <out.md>     # `ncu --metrics gpu__time_duration.sum --csv` list
    python tools/ncu_summary.py full <raw.csv> <out.md>              # `ncu -i x.ncu-rep --page raw --csv` of a --set full capture

The launch list is reduced to one row per kernel (launches, mean/min/max duration, share of the total);
the full capture to the handful of metrics the roofline argument uses.
"""
import collections
import csv
import sys

FULL_METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe insts"),
    ("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
]


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.split("(")[0]


def launches(src, dst):
    hdr, agg, order = None, collections.OrderedDict(), []
    for r in csv.reader(open(src, errors="replace")):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        if d.get("Metric Unit") == "us":
            v *= 1e3
        agg.setdefault(short(d["Kernel Name"]), []).append(v)
    total = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("| kernel | launches | mean us | min us | max us | share of GPU time |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("| `%s` | %d | %.2f | %.2f | %.2f | %.1f %% |\n" %
                    (k, len(v), sum(v) / len(v) / 1e3, min(v) / 1e3, max(v) / 1e3, 100 * sum(v) / total))
        f.write("\ntotal GPU time in the capture: %.3f ms over %d launches "
                "(per-launch times under ncu are serialised and cold-cache: read the SHARES)\n" %
                (total / 1e6, sum(len(v) for v in agg.values())))


def full(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = [(hdr.index(m), lab) for m, lab in FULL_METRICS if m in hdr]
    with open(dst, "w") as f:
        f.write("| kernel | " + " | ".join(lab for _, lab in cols) + " |\n|---|" + "---|" * len(cols) + "\n")
        for r in rows[2:]:
            f.write("| `%s` | " % short(r[ki]) + " | ".join("%s %s" % (r[i], units[i]) for i, _ in cols) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
