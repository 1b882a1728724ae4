"""Summarise an ncu launch list (csv written by tools/gpu_round.sh: gpu__time_duration.sum [+ dram bytes] per launch of
one bench.py step) into profiles/launches_<tag>.md and profiles/traffic.json.   python tools/summarize_ncu.py <csv> <tag>"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def short(name):
    name = re.sub(r"\(CUtensorMap_st.*", "", name)
    name = re.sub(r"\((const )?(__half|float|long long|int|void).*", "", name)
    return name.replace("void ", "").replace("db1::", "").strip()[:70]


def main(path, tag):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kid, kn, mn, mv, mu = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(",", ""))
        except ValueError:
            continue
        unit = r[mu].lower()
        if "duration" in r[mn]:
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(unit, 1e-3)
        elif "bytes" in r[mn]:
            v *= {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
        per.setdefault(r[kid], {"name": short(r[kn])})[r[mn]] = v
    agg = collections.defaultdict(lambda: dict(n=0, us=0.0, rd=0.0, wr=0.0))
    for d in per.values():
        a = agg[d["name"]]
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["us"] for a in agg.values())
    lines = ["# ncu launch list, %s" % tag, "",
             "Source: `%s` (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`"
             " around `python bench.py --steps 1 --warmup 1 --profile-only`, i.e. 2 forward+backward steps of DB1-1.3B B=4 L=1024)." % os.path.basename(path),
             "Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's CUDA-event breakdown, not absolutes.", "",
             "| kernel | launches | total us | share | avg us | avg DRAM read MB | avg DRAM write MB |", "|---|---:|---:|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        lines.append("| `%s` | %d | %.1f | %.1f%% | %.1f | %.1f | %.1f |" % (k, a["n"], a["us"], 100 * a["us"] / tot, a["us"] / a["n"],
                                                                      a["rd"] / a["n"] / 1e6, a["wr"] / a["n"] / 1e6))
    lines.append("")
    lines.append("total %.1f us over %d launches" % (tot, sum(a["n"] for a in agg.values())))
    out = os.path.join(ROOT, "profiles", "launches_%s.md" % tag)
    open(out, "w").write("\n".join(lines) + "\n")
    g = [a for k, a in agg.items() if k.startswith("gemm_kernel")]
    if g and sum(a["rd"] + a["wr"] for a in g) > 0:
        n = sum(a["n"] for a in g)
        json.dump({"gemm_dram_bytes_per_launch": sum(a["rd"] + a["wr"] for a in g) / n, "gemm_launches": n, "source": os.path.basename(out)},
                  open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
