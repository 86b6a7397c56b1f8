"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md: tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, mbarrier -> SYNCS, legacy mma.sync -> HMMA).
Runs on the CPU box: cuobjdump -sass on the shipped library.  Usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "salience-detr_b200", "lib", "libsdetr_b200.so")
PATTERNS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "SYNCS", "HMMA", "LDGSTS",
            "REDG", "RED.E", "ATOMG", "LDG.E.128", "LDS.128", "SHFL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    counts, order, cur, i = collections.OrderedDict(), [], None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[i] if i < len(names) else m.group(1)
            i += 1
            cur = re.sub(r"\(.*$", "", cur.replace("(anonymous namespace)", "{anon}"))
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["instr"] += 1
            for p in PATTERNS:
                if op.startswith(p):
                    counts[cur][p] += 1
    cols = [p for p in PATTERNS if any(c[p] for c in counts.values())]
    print("# SASS evidence per kernel of salience-detr_b200/lib/libsdetr_b200.so (cuobjdump -sass; tools/sass_summary.py)")
    print("# kernel".ljust(70) + "".join(c.rjust(10) for c in ["instr"] + cols))
    for k, c in counts.items():
        print(k[:69].ljust(70) + "".join(str(c[x]).rjust(10) for x in ["instr"] + cols))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("TOTAL".ljust(70) + "".join(str(tot[x]).rjust(10) for x in ["instr"] + cols))


if __name__ == "__main__":
    main()
