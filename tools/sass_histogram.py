#!/usr/bin/env python3
"""Per-kernel SASS opcode histogram of libnbody_b200.so (cuobjdump -sass; runs here, no GPU needed):

    python tools/sass_histogram.py > profiles/rNN_sass_opcodes.md

One row per kernel of the library with the counts of the instructions that prove which hardware paths the code uses:
UBLKCP (1-D bulk TMA), SYNCS (mbarrier), STAS (st.async to distributed shared memory), FFMA2 / FADD2 / FMUL2 (packed
f32x2), DFMA / DADD / DMUL, MUFU.RSQ / MUFU.RSQ64H, plus the register count and spill bytes from lib/ptxas.log."""
import collections
import hashlib
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "nbody-gnn-hpc_b200" / "lib" / "libnbody_b200.so"
COLS = ["UBLKCP", "SYNCS", "STAS", "FFMA2", "FADD2", "FMUL2", "FFMA", "DFMA", "DADD", "DMUL", "MUFU.RSQ", "MUFU.RSQ64H",
        "LDS", "LDG", "STG", "ATOMG", "RED", "BAR", "UCGABAR_ARV", "ACQBULK"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    return dict(zip(names, out))


def source_hash() -> str:
    h = hashlib.sha256()
    for f in sorted((ROOT / "nbody-gnn-hpc_b200" / "csrc").glob("*.cu*")) + [ROOT / "include" / "nbody_b200.h"]:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    return h.hexdigest()[:16]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], stdout=subprocess.PIPE, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    arch = set(re.findall(r"arch = (sm_\w+)", sass))
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            base = op.split(".")[0]
            cur[base] += 1
            if op.startswith("MUFU.RSQ64H"):
                cur["MUFU.RSQ64H"] += 1
            elif op.startswith("MUFU.RSQ"):
                cur["MUFU.RSQ"] += 1
    regs = {}
    log = (LIB.parent / "ptxas.log").read_text() if (LIB.parent / "ptxas.log").exists() else ""
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'(.*?)Used (\d+) registers", log, flags=re.S):
        sp = re.search(r"(\d+) bytes spill stores", m.group(2))
        regs[m.group(1)] = (int(m.group(3)), int(sp.group(1)) if sp else 0)
    names = demangle(list(kernels))
    print(f"# SASS opcode histogram of libnbody_b200.so\n\narch: {', '.join(sorted(arch))}; CUDA sources hash "
          f"`{source_hash()}` (the hash bench.py reports as `source_hash`); `python tools/sass_histogram.py`\n")
    print("| kernel | regs | spill B | instr | " + " | ".join(COLS) + " |")
    print("|---|---|---|---|" + "---|" * len(COLS))
    for mangled, c in kernels.items():
        name = names[mangled]
        name = re.sub(r"\(.*", "", name).replace("void nb::", "").replace("nb::", "")
        r = regs.get(mangled, ("", ""))
        print(f"| `{name}` | {r[0]} | {r[1]} | {c['_total']} | " + " | ".join(str(c.get(k, 0) or "") for k in COLS) + " |")


if __name__ == "__main__":
    sys.exit(main())
