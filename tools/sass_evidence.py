"""SASS evidence per kernel of libsketchedit_b200.so: counts of the Blackwell-native mnemonics (B200_PROFILING.md):
UTCHMMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UBLKCP (TMA / bulk copies), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), HMMA (legacy mma.sync: must be 0).  ->  markdown on stdout

    python tools/sass_evidence.py > profiles/r02_sass_mnemonics.md
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "sketchedit_b200", "libsketchedit_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
MN = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "R2UR"]
cur, counts = None, collections.OrderedDict()
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        for k in MN:
            if k == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    counts[cur][k] += 1
            elif op.split(".")[0] == k:
                counts[cur][k] += 1
dem = subprocess.run(["cu++filt"] + list(counts), stdout=subprocess.PIPE, text=True).stdout.splitlines()
print("# SASS mnemonics per kernel (`cuobjdump -sass sketchedit_b200/libsketchedit_b200.so`)\n")
print("UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,")
print("UTCBAR = tcgen05.commit, SYNCS = mbarrier ops. HMMA (legacy mma.sync) must be absent.\n")
print("| kernel | " + " | ".join(MN) + " |")
print("|---|" + "---:|" * len(MN))
for (k, c), d in zip(counts.items(), dem):
    if not any(c[x] for x in MN if x != "R2UR"):
        continue
    name = d.replace("(int)", "")
    name = (name.split(">(")[0] + ">") if ">(" in name else re.sub(r"\(.*", "", name)
    name = name.replace("void se::", "").replace("se::", "")
    print("| `%s` | " % name + " | ".join(str(c[x]) for x in MN) + " |")
