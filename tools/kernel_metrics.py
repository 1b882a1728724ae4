"""Summarise an ncu metrics pass over one bench step (long-format csv written by tools/gpu_round.sh `kmetrics`) into
profiles/kernels_<tag>.md: one row per kernel instantiation the step launches, averaged over its launches.
    python tools/kernel_metrics.py <csv> [<csv> ...] <tag>"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("gpu__time_duration.sum", "avg us", 1e-3), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("dram__bytes_read.sum", "DRAM rd MB", 1e-6), ("dram__bytes_write.sum", "DRAM wr MB", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts", 1),
        ("launch__registers_per_thread", "regs", 1), ("launch__grid_size", "grid", 1)]
UNIT = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
        "byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}


def short(name):
    name = re.sub(r"\(CUtensorMap_st.*", "", name)
    name = re.sub(r"\((const )?(__half|float|long long|int|void|unsigned|db1::|BandParams|DecParams|KvParams).*", "", name)
    return name.replace("void ", "").replace("db1::", "").strip()[:84]


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    kid, kn, mn, mv, mu = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        v *= UNIT.get(r[mu].lower(), 1.0)
        per.setdefault((path, r[kid]), {"name": short(r[kn])})[r[mn]] = v
    return per


def main(paths, tag):
    lines = ["# Per-kernel ncu metrics, %s" % tag, "",
             "One `ncu --metrics ... --clock-control none` pass per workload (tools/gpu_round.sh kmetrics): every launch of the profiled "
             "steps, averaged per kernel instantiation. Times under ncu are cold-cache and serialised (compare shares, not "
             "absolutes, with bench.py's CUDA-event numbers); tensor % / issue % are of the peak over active cycles; DRAM % and L2 % "
             "of peak over elapsed cycles.", ""]
    for path in paths:
        per = load(path)
        agg = collections.OrderedDict()
        for d in per.values():
            a = agg.setdefault(d["name"], collections.defaultdict(float))
            a["n"] += 1
            for k, _l, _s in COLS:
                a[k] += d.get(k, 0.0)
        tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
        lines += ["## %s" % os.path.basename(path), "",
                  "| kernel | launches | share | " + " | ".join(c[1] for c in COLS) + " |", "|---|---:|---:|" + "---:|" * len(COLS)]
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
            vals = []
            for key, _l, sc in COLS:
                v = a[key] / a["n"] * sc
                vals.append("%.1f" % v if v < 1e5 else "%.3g" % v)
            lines.append("| `%s` | %d | %.1f%% | %s |" % (k, a["n"], 100 * a["gpu__time_duration.sum"] / tot, " | ".join(vals)))
        lines += ["", "total %.1f us over %d launches" % (tot / 1e3, sum(a["n"] for a in agg.values())), ""]
    out = os.path.join(ROOT, "profiles", "kernels_%s.md" % tag)
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main(sys.argv[1:-1], sys.argv[-1])
