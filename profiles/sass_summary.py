"""Per-kernel SASS evidence of the Blackwell-native paths (B200_PROFILING.md "What proves a Blackwell-native kernel"):
counts of tcgen05.mma (UTC*MMA), tcgen05.ld/st (LDTM/STTM), TMA (UTMALDG/UTMASTG/UBLKCP) and, as the thing that must
NOT be there, legacy mma.sync (HMMA) in every kernel of libb2retr.so.
    python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

so = Path(__file__).resolve().parent.parent / "movie_recommender_demo_b200" / "libb2retr.so"
out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
pats = {"UTC*MMA (tcgen05.mma)": r"\bUTC[A-Z]*MMA\b", "LDTM (tcgen05.ld)": r"\bLDTM\b", "STTM (tcgen05.st)": r"\bSTTM\b",
        "UTMALDG (TMA load)": r"\bUTMALDG\b", "UTMASTG (TMA store)": r"\bUTMASTG\b", "UBLKCP (bulk copy)": r"\bUBLKCP\b",
        "SYNCS (mbarrier)": r"\bSYNCS\b", "LDGSTS (cp.async)": r"\bLDGSTS\b", "HMMA (legacy mma.sync)": r"\bHMMA\b"}
kern, counts, size = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "").replace("b2r::", "")
        kern = re.sub(r"\((?:CUtensorMap|b2r|float|int|long|unsigned|const|__nv|uint|bool|char|__half|void).*", "", kern)
        counts[kern] = collections.Counter()
        size[kern] = 0
        continue
    if kern and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
        size[kern] += 1
        for name, pat in pats.items():
            if re.search(pat, line):
                counts[kern][name] += 1
print(f"# cuobjdump -sass {so.name}: instruction counts per kernel (sm_100a)")
cols = list(pats)
print("kernel".ljust(64) + "".join(c.split(" ")[0].rjust(9) for c in cols) + "   SASS instrs")
tot = collections.Counter()
for k, c in counts.items():
    print(k[:63].ljust(64) + "".join(str(c.get(n, 0) or "-").rjust(9) for n in cols) + str(size[k]).rjust(14))
    tot.update(c)
print("TOTAL".ljust(64) + "".join(str(tot.get(n, 0)).rjust(9) for n in cols))
print("\nlegend: " + "; ".join(cols))
