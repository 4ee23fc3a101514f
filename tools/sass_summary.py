"""SASS instruction summary of libtetris_b200.so per kernel (cuobjdump -sass, read in the build container; no GPU needed).

    python tools/sass_summary.py profiles/r02_sass_summary.md

Per kernel: instruction count and the mnemonics that identify the hardware paths used -- UBLKCP (1-D bulk TMA), SYNCS / ATOMS
(mbarrier), ACQBULK (programmatic dependent launch), LDGSTS (cp.async), IDP (dp4a), VABSDIFF4 / VIMNMX (packed bytes), PRMT,
LOP3, SHF, POPC, BAR, ... and the absence of tensor-core instructions (UTCMMA / LDTM / HMMA: no contraction on this path)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tetris_gymnasium_b200", "libtetris_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary.md")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip() or s  # noqa: E731
kernels, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[T\d]+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur is not None:
        cur[m.group(1)] += 1
KEY = ["UBLKCP", "SYNCS", "ACQBULK", "LDGSTS", "IDP", "VABSDIFF4", "VIMNMX", "PRMT", "LOP3", "SHF", "POPC", "FLO", "BREV", "BAR", "LDS", "STS", "LDG", "STG",
       "ATOMS", "ATOMG", "RED", "UTCMMA", "LDTM", "HMMA", "IMMA"]
with open(out, "w") as f:
    f.write("# SASS instruction summary of `libtetris_b200.so` (sm_100a), `cuobjdump -sass`\n\n"
            "Counts are static instructions per kernel.  `UBLKCP` = 1-D bulk TMA copy (`cp.async.bulk`), `SYNCS` = mbarrier, `ACQBULK` = programmatic "
            "dependent launch (`griddepcontrol`), `LDGSTS` = `cp.async`, `IDP` = `dp4a`, `VABSDIFF4` = packed-byte absolute differences.  "
            "No `UTCMMA` / `LDTM` / `HMMA`: the path has no dense contraction, tensor cores are deliberately unused.\n\n")
    f.write("| kernel | instr | " + " | ".join(KEY) + " |\n|---|---|" + "---|" * len(KEY) + "\n")
    tot = collections.Counter()
    for name, c in kernels.items():
        d = demangle(name)
        d = re.sub(r"\((?:int|bool)\)", "", re.sub(r"^void ", "", d))
        d = d[:d.rfind(">(") + 1] if ">(" in d else d.split("(")[0]
        f.write(f"| `{d[:70]}` | {sum(c.values())} | " + " | ".join(str(c.get(k, 0)) for k in KEY) + " |\n")
        tot.update(c)
    f.write(f"| **all kernels** | {sum(tot.values())} | " + " | ".join(str(tot.get(k, 0)) for k in KEY) + " |\n")
print(open(out).read()[:3000])
