"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel, from the shipped library:
UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAPF = TMA load / store / prefetch,
HMMA = legacy mma.sync (must be 0).     python tools/sass_opcodes.py > profiles/sass_opcodes_<round>.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "bdm-db1_b200", "libdb1_sm100.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMAREDG", "SYNCS", "HMMA", "REDG", "RED", "ATOMG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            per[cur]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    per[cur][o] += 1
    dem = subprocess.run(["c++filt"], input="\n".join(per.keys()), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode counts per kernel of libdb1_sm100.so (cuobjdump -sass; tools/sass_opcodes.py)")
    print("# %-88s %7s " % ("kernel", "instrs") + " ".join("%8s" % o for o in OPS))
    tot = collections.Counter()
    for (name, c), d in zip(per.items(), dem):
        d = re.sub(r"\(.*", "", d).replace("void ", "").replace("db1::", "")
        print("%-90s %7d " % (d[:90], c["_total"]) + " ".join("%8d" % c[o] for o in OPS))
        tot.update(c)
    print("%-90s %7d " % ("TOTAL", tot["_total"]) + " ".join("%8d" % tot[o] for o in OPS))
    if tot["HMMA"]:
        sys.exit("legacy HMMA present")


if __name__ == "__main__":
    main()
