"""profiles/sass_summary.md: per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA (B200_PROFILING.md) in
the product library, from `cuobjdump -sass`. usage: python tools/sass_summary.py [lib] > profiles/sass_summary.md"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "pixeltable_yolox_b200" / "libyx_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
MN = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "UTCATOMSWS", "SYNCS", "MUFU", "STG", "LDG", "STS", "LDS", "ACQBULK", "UCGABAR", "CCTL"]
per = collections.OrderedDict()
arch, cur = None, None
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch = m.group(1)
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = per.setdefault(m.group(1), collections.Counter())
        cur["arch:" + str(arch)] += 1
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        cur["_total"] += 1
        for k in MN:
            if op.startswith(k):
                cur[k] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# SASS summary of `{Path(lib).name}` (cuobjdump -sass, {len(per)} kernels)\n")
print("Mnemonics: `UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld (TMEM -> registers), `UTMALDG` = TMA tensor load, `UBLKCP` = bulk copy,")
print("`UTCBAR` = tcgen05.commit -> mbarrier, `SYNCS` = mbarrier ops, `MUFU` = special-function unit, `UCGABAR` = cluster barrier.\n")
print("| kernel | arch | SASS instr | " + " | ".join(MN) + " |")
print("|---|---|---:|" + "---:|" * len(MN))
tot = collections.Counter()
for (name, c), dn in zip(per.items(), demangle):
    short = re.sub(r"\(.*", "", dn).replace("void ", "").replace("yx::", "")
    a = ",".join(k[5:] for k in c if k.startswith("arch:"))
    print(f"| `{short}` | {a} | {c['_total']} | " + " | ".join(str(c[k]) if c[k] else "" for k in MN) + " |")
    tot.update({k: c[k] for k in MN}); tot["_total"] += c["_total"]
print(f"| **total** | | {tot['_total']} | " + " | ".join(str(tot[k]) for k in MN) + " |")
