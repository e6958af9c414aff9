#!/usr/bin/env python3
"""Per-kernel SASS evidence of what the library is made of: which kernels use the TMA engine (UBLKCP = 1-D cp.async.bulk),
mbarriers (SYNCS), warp reductions (REDUX), shared / global atomics (ATOMS, ATOMG, RED), packed 16-bit integer ops
(VIADD.16x2, VIMNMX.U16x2 ...) and float64 arithmetic (DFMA, DMUL, DADD), and that nothing else hides in the binary
(no UTMALDG / tensor-core UTC*MMA is expected: the path streams 1-D runs and has no contraction; sm_100a cubins only).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "repas_vision_b200", "librepasvision.so")
COLS = ["UBLKCP", "UTMALDG", "UTCMMA", "SYNCS", "BAR", "REDUX", "ATOMS", "ATOMG", "RED", "x16x2", "DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS"]


def classify(op):
    if op.startswith("UBLKCP"):
        return "UBLKCP"
    if op.startswith("UTMALDG") or op.startswith("UTMASTG"):
        return "UTMALDG"
    if re.match(r"UTC.*MMA", op):
        return "UTCMMA"
    if op.startswith("SYNCS"):
        return "SYNCS"
    if op.startswith("BAR"):
        return "BAR"
    if op.startswith("REDUX"):
        return "REDUX"
    if op.startswith("ATOMS"):
        return "ATOMS"
    if op.startswith("ATOMG") or op.startswith("ATOM."):
        return "ATOMG"
    if op.startswith("RED."):
        return "RED"
    if "16x2" in op:
        return "x16x2"
    for k in ("DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS"):
        if op.startswith(k):
            return k
    return None


def main():
    elf = subprocess.run(["cuobjdump", "--list-elf", SO], capture_output=True, text=True).stdout
    archs = sorted(set(re.findall(r"sm_(\d+a?)", elf)))
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    fam = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("void ", "")
            base = re.sub(r"[<(].*", "", name).split("::")[-1]
            cur = fam.setdefault(base, {"variants": 0, "instr": 0, "ops": collections.Counter()})
            cur["variants"] += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.x]*)", line)
        if m and cur is not None:
            cur["instr"] += 1
            c = classify(m.group(1))
            if c:
                cur["ops"][c] += 1
    print(f"# {os.path.relpath(SO, ROOT)}: cubin architectures {archs}; instruction counts are static SASS lines summed over the")
    print("# template instantiations of each kernel (variants); x16x2 = packed 16-bit integer ops (VIADD.16x2, VIMNMX.U16x2, ...)")
    print("kernel".ljust(24) + "variants".rjust(9) + "instr".rjust(9) + "".join(c.rjust(8) for c in COLS))
    tot = collections.Counter()
    for k, v in fam.items():
        print(k[:23].ljust(24) + str(v["variants"]).rjust(9) + str(v["instr"]).rjust(9) + "".join(str(v["ops"].get(c, 0)).rjust(8) for c in COLS))
        tot.update(v["ops"])
    print("TOTAL".ljust(24) + str(sum(v["variants"] for v in fam.values())).rjust(9) + str(sum(v["instr"] for v in fam.values())).rjust(9)
          + "".join(str(tot.get(c, 0)).rjust(8) for c in COLS))


if __name__ == "__main__":
    main()
