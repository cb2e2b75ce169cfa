"""Per-kernel SASS census of thunder_b200/lib/libthunder_b200.so (cuobjdump -sass): the instructions that identify how each kernel
moves its bytes.  usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "thunder_b200" / "lib" / "libthunder_b200.so"
COLS = [("LDG.256", r"\bLDG\.E(\.\w+)*\.256"), ("LDG.128", r"\bLDG\.E(\.\w+)*\.128"), ("LDG.64", r"\bLDG\.E(\.\w+)*\.64\b"),
        ("LDS.128", r"\bLDS\.128"), ("REDG.F32x4", r"\bREDG\.E\.ADD\.F32x4"), ("REDG.F64", r"\b(REDG|RED)\.E\.ADD\.F64"),
        ("ATOMG", r"\bATOMG"), ("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("DFMA", r"\bDFMA"), ("FFMA", r"\bFFMA"),
        ("MUFU", r"\bMUFU"), ("SHFL", r"\bSHFL"), ("VOTE", r"\bVOTE"), ("NANOSLEEP", r"\bNANOSLEEP"), ("HMMA/UTCMMA", r"\b(HMMA|UTC\w*MMA|QGMMA)")]


def main():
    out = subprocess.check_output(["cuobjdump", "-sass", str(LIB)], text=True)
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.check_output(["c++filt", m.group(1)], text=True).strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            cur = name
            k = 2
            while cur in counts:
                cur = f"{name} #{k}"; k += 1
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        for col, pat in COLS:
            if re.search(pat, line):
                counts[cur][col] += 1
    print(f"SASS summary of thunder_b200/lib/libthunder_b200.so (cubins: {', '.join(arch)}; `cuobjdump -sass`; tools/sass_summary.py): static instruction")
    print("counts per kernel.  LDG.256 = the 256-bit gathers of the cell / quad volume layouts (sm_100-only width); REDG.F32x4 = the 16-byte vector")
    print("reductions of the M kernels (red.global.add.v4.f32); REDG.F64 = double atomics; NANOSLEEP = the bounded spin of the lockstep barrier")
    print("(expect_multi_kernel); VOTE = the ballots of the warp-parallel polar method (particle filter); UBLKCP / SYNCS = TMA bulk copies + mbarrier,")
    print("only in the staged alternative E kernel (expect_impl 2).  No tensor-core instructions (HMMA / UTC*MMA) anywhere: no dense contraction on")
    print("this path is allowed onto them (DESIGN.md section 4.12).\n")
    w = max(len(k) for k in counts) + 2
    print("kernel".ljust(w) + "".join(c.rjust(12) for c, _ in COLS))
    for k in sorted(counts):
        print(k.ljust(w) + "".join(str(counts[k][c]).rjust(12) for c, _ in COLS))


if __name__ == "__main__":
    sys.exit(main())
