#!/usr/bin/env python
"""sass_count.py -- instruction mix of the hot loops of libsatmc.so, read off the SASS.

    python tools/sass_count.py [--lib PATH] [--kernel SUBSTR] [--json OUT] [--dump OUT.txt]

For every kernel whose (demangled) name contains SUBSTR (default: the fused counting kernel
`k_count<satmc::DirectSrc, false, false, false>`) the script disassembles the function with `cuobjdump -sass`,
finds the loops (a backward branch closes a loop: [target, branch]) and reports, for every innermost
loop with at least --min-instr instructions: total instructions, and counts of IMAD.WIDE, other integer
multiply-adds, FP32 (FFMA/FMUL/FADD/FMNMX/FSETP/FSEL), MUFU, LOP3/shift/PRMT, I2F/F2I conversions, loads /
stores (local-memory ones flagged: a spill inside a hot loop shows up as LDL/STL), branches and the rest.

bench.py reads the JSON this writes (profiles/r2_sass_k_count_hotloop.json) for the per-test instruction
constants of its roofline line, so the numbers in the bench line can be re-derived from a committed artefact:
    instr_per_group           -> plain issue slots (what ncu's smsp__issue_active counts)
    imad_wide_per_group       -> the quarter-rate instruction that binds the fused loop (tools/ubench.cu)
No GPU is needed: this reads the cubin embedded in the shared library.
"""
from __future__ import annotations

import argparse
import collections
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_LIB = os.path.join(ROOT, "convex-2d-gpu-collision-detection_b200", "libsatmc.so")
CUOBJDUMP = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")

INSTR_RE = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);")


def classify(op: str, rest: str) -> str:
    base = op.split(".")[0]
    if base == "IMAD" and ".WIDE" in op:
        return "IMAD.WIDE"
    if base == "IMAD" and ".HI" in op:
        return "IMAD.HI"
    if op.startswith("IMAD.MOV") or op.startswith("IMAD.SHL") or op.startswith("IMAD.IADD"):
        return "IMAD(mov/shl/add)"
    if base in ("IMAD", "IMUL"):
        return "IMAD"
    if base in ("FFMA2", "FMUL2", "FADD2"):
        return "FP32x2"
    if base in ("FFMA", "FMUL", "FADD", "FMNMX", "FMNMX3", "FSETP", "FSEL", "FSET", "FCHK"):
        return "FP32"
    if base == "MUFU":
        return "MUFU"
    if base in ("LOP3", "LOP", "SHF", "PRMT", "LEA", "IADD3", "IADD", "ISETP", "SEL", "VIADD", "IABS", "PLOP3", "POPC", "FLO",
                "BREV", "SGXT", "BMSK", "VIMNMX", "VIMNMX3", "IMNMX", "P2R", "R2P"):
        return "ALU(int/logic)"
    if base in ("I2F", "I2FP", "F2I", "F2F", "F2FP", "I2I", "FRND"):
        return "CVT"
    if base in ("LDL", "STL"):
        return "LOCAL(spill)"
    if base in ("LDG", "STG", "LDS", "STS", "LDC", "LD", "ST", "ATOMG", "ATOMS", "RED", "ATOM", "LDSM", "UTMALDG", "SYNCS"):
        return "MEM"
    if base in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "EXIT", "WARPSYNC", "BAR", "NOP", "YIELD", "BREAK", "BMOV", "JMP"):
        return "CTRL"
    if base in ("HFMA2", "MOV", "CS2R", "S2R", "S2UR", "R2UR", "UMOV", "SHFL", "VOTE", "VOTEU", "REDUX", "MATCH"):
        return "MOVE/WARP"
    if base.startswith("U") or base == "LDCU":
        return "UNIFORM"
    return "OTHER:" + base


def disassemble(lib: str):
    """yields (mangled name, [(addr, opcode, rest)])"""
    out = subprocess.run([CUOBJDUMP, "-sass", lib], check=True, capture_output=True, text=True).stdout
    name, instrs = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name is not None:
                yield name, instrs
            name, instrs = m.group(1), []
            continue
        m = INSTR_RE.match(line)
        if m and name is not None:
            instrs.append((int(m.group(1), 16), m.group(2), m.group(3)))
    if name is not None:
        yield name, instrs


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout
        return dict(zip(names, out.splitlines()))
    except (OSError, subprocess.CalledProcessError):
        return {n: n for n in names}


def loops_of(instrs):
    """[(start_index, end_index)] of loops closed by a backward BRA, innermost first"""
    addr_to_idx = {a: i for i, (a, _, _) in enumerate(instrs)}
    found = []
    for i, (a, op, rest) in enumerate(instrs):
        if op.split(".")[0] != "BRA":
            continue
        m = re.search(r"0x([0-9a-f]+)", rest)
        if not m:
            continue
        t = int(m.group(1), 16)
        if t <= a and t in addr_to_idx:
            found.append((addr_to_idx[t], i))
    found.sort(key=lambda se: se[1] - se[0])
    return found


def summarise(instrs, s, e):
    body = instrs[s:e + 1]
    cnt = collections.Counter(classify(op, rest) for _, op, rest in body)
    ops = collections.Counter(op.split(".")[0] for _, op, _ in body)
    text = "\n".join(op + re.sub(r"0x[0-9a-f]+", "X", rest) for _, op, rest in body)       # schedule and register assignment
    return {"start": hex(body[0][0]), "end": hex(body[-1][0]), "instr": len(body), "classes": dict(sorted(cnt.items())),
            "fingerprint": hashlib.sha1(text.encode()).hexdigest()[:16],
            "imad_wide": cnt.get("IMAD.WIDE", 0), "fp32": cnt.get("FP32", 0), "fp32x2": cnt.get("FP32x2", 0), "mufu": cnt.get("MUFU", 0),
            "local_spill": cnt.get("LOCAL(spill)", 0),
            "opcodes": dict(sorted(ops.items(), key=lambda kv: -kv[1]))}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--lib", default=DEFAULT_LIB)
    ap.add_argument("--kernel", default="k_count<satmc::DirectSrc, false, false, false>")
    ap.add_argument("--min-instr", type=int, default=60)
    ap.add_argument("--json", default=None)
    ap.add_argument("--dump", default=None, help="write the SASS of the reported loops here")
    args = ap.parse_args()
    funcs = list(disassemble(args.lib))
    names = demangle([n for n, _ in funcs])
    report, dump = {}, []
    for mangled, instrs in funcs:
        nice = names[mangled]
        if args.kernel not in nice:
            continue
        loops = loops_of(instrs)
        inner = []
        for s, e in loops:                                  # innermost = contains no other reported loop
            if e - s + 1 < args.min_instr:
                continue
            if any(s <= s2 and e2 <= e and (s2, e2) != (s, e) and e2 - s2 + 1 >= args.min_instr for s2, e2 in loops):
                continue
            inner.append((s, e))
        entry = {"function_instr": len(instrs), "loops": [summarise(instrs, s, e) for s, e in sorted(inner)]}
        report[nice] = entry
        print(f"== {nice}   ({len(instrs)} instructions)")
        for (s, e), L in zip(sorted(inner), entry["loops"]):
            print(f"   loop {L['start']}..{L['end']}: {L['instr']} instr | IMAD.WIDE {L['imad_wide']} | FP32 {L['fp32']} + {L['fp32x2']} packed | MUFU {L['mufu']} | "
                  f"spill ld/st {L['local_spill']}")
            print("      " + ", ".join(f"{k} {v}" for k, v in L["classes"].items()))
            if args.dump:
                dump.append(f"// {nice}  loop {L['start']}..{L['end']}  ({L['instr']} instructions)")
                dump += [f"        /*{a:04x}*/  {op}{rest} ;" for a, op, rest in instrs[s:e + 1]]
                dump.append("")
    if not report:
        print(f"no kernel matching {args.kernel!r} in {args.lib}", file=sys.stderr)
        return 1
    if args.json:
        with open(args.json, "w") as f:
            json.dump({"lib": os.path.relpath(args.lib, ROOT), "kernel_filter": args.kernel, "kernels": report}, f, indent=1)
    if args.dump:
        with open(args.dump, "w") as f:
            f.write("\n".join(dump))
    return 0


if __name__ == "__main__":
    sys.exit(main())
