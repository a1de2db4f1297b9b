#!/usr/bin/env python
"""Regenerates profiles/r2_sass_mnemonics.md: per kernel of the sm_100a objects under scann-rust_b200/lib/obj, the counts
of the Blackwell-specific / hot-path SASS mnemonics (cuobjdump -sass | c++filt).  Run after scann-rust_b200/build.py."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "scann-rust_b200", "lib", "obj")
KEEP = re.compile(r"^(UTC\w+(\.\w+)*|UTMA\w+(\.\w+)*|LDTM\.?|STTM\.?|SYNCS(\.\w+)+|ELECT|R2UR(\.BROADCAST)?|BRA\.U\.ANY|PRMT|"
                  r"IDP\.2A(\.\w+)+|LDG\.E\.128(\.CONSTANT)?|ATOMS(\.\w+)*|FMNMX3)$")
OBJECTS = ["tcscan.o", "tc_gemm.o", "treeah.o"]
HEAD = """# SASS evidence (cuobjdump -sass of the sm_100a objects built from this tree; generator: tools/sass_summary.py)

Counts of the Blackwell-specific / hot-path mnemonics per kernel: UTCIMMA = tcgen05.mma kind::i8, UTCHMMA = tcgen05.mma kind::f16,
UTMALDG = cp.async.bulk.tensor (TMA load), LDTM / STTM = tcgen05.ld / tcgen05.st, UTCBAR = tcgen05.commit, SYNCS = mbarrier,
BRA.U.ANY = a waterfall loop around a uniform-datapath instruction (0 around the MMAs since the issue loops run under elect.sync),
PRMT / IDP = the register-LUT lookup and IDP.2A accumulation of the register-LUT scan.  `instr` = SASS instructions of the kernel.
"""


def main():
    out = [HEAD]
    for o in OBJECTS:
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(OBJ, o)], capture_output=True, text=True, check=True).stdout
        name, counts, n = None, None, 0

        def emit():
            if name is None:
                return
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
            body = ", ".join(f"{k} x{v}" for k, v in sorted(counts.items()))
            out.append(f"### {o}  {dem[:150]}\ninstr {n}" + (", " + body if body else "") + "\n")
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                emit()
                name, counts, n = m.group(1), collections.Counter(), 0
                continue
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if m and name is not None:
                n += 1
                op = m.group(1)
                if KEEP.match(op):
                    counts[op] += 1
        emit()
    path = os.path.join(ROOT, "profiles", "r2_sass_mnemonics.md")
    with open(path, "w") as f:
        f.write("\n".join(out))
    print(path, len(out) - 1, "kernels")


if __name__ == "__main__":
    sys.exit(main())
