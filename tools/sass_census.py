#!/usr/bin/env python3
"""Static census of the simulation kernels' SASS: for every kernel of libhw1f.so, the loop body that
carries the Box-Muller work (the innermost backward branch whose body holds the most MUFU instructions)
is counted by pipe -- MUFU (XU), packed FP32x2 (FFMA2/FADD2/FMUL2: two dispatch cycles each), scalar FP32,
integer/logic (ALU), conversions, shared/global memory, shuffles, everything else -- and the resource demand
of one body per warp is derived:

    dispatch cycles = instructions + packed instructions        (a packed FP32x2 holds the port 2 cycles)
    XU cycles       = 8 x MUFU (+ 4 x I2FP/F2I... listed, not added: they issue beside MUFU)

    python tools/sass_census.py [lib.so] > profiles/r02_sass_census.json

Needs only cuobjdump (no GPU).  The roofline argument of DESIGN.md section 4 ("299 dispatch / 320 XU cycles
per loop body of 40 path-steps") is reproduced from this file.
"""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "monte-carlo-simulation-of-hull-white-model-and-sensitivities-computation_b200", "lib",
                   "libhw1f.so")

INSTR = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*")
PACKED = ("FFMA2", "FADD2", "FMUL2")
FP32 = ("FFMA", "FADD", "FMUL", "FMNMX", "FSEL", "FSETP", "FSET", "FCHK")
ALU = ("LOP3", "SHF", "IADD3", "IADD", "IMAD", "LEA", "ISETP", "SEL", "MOV", "PRMT", "IABS", "LOP", "SHL", "SHR",
       "PLOP3", "IMNMX", "VIADD", "VIMNMX", "UIADD3", "ULOP3", "UMOV", "USHF", "UIMAD", "ULEA", "UISETP", "USEL")
CONV = ("I2FP", "I2F", "F2I", "F2F", "F2FP")
MEM = ("LDS", "STS", "LDG", "STG", "LDC", "LDCU", "LD", "ST", "ATOMS", "ATOMG", "RED", "LDSM", "ULDC")
SHFL = ("SHFL", "VOTE", "REDUX", "MATCH")


def classify(op):
    base = op.split(".")[0]
    if base == "MUFU":
        return "mufu"
    if base in PACKED:
        return "packed_fp32x2"
    if base in CONV:
        return "convert"
    if base in FP32:
        return "fp32"
    if base in SHFL:
        return "shuffle_vote"
    if base in MEM:
        return "memory"
    if base in ("DADD", "DMUL", "DFMA", "DSETP"):
        return "fp64"
    if base in ("BRA", "BSSY", "BSYNC", "EXIT", "BAR", "WARPSYNC", "CALL", "RET", "NOP", "BREAK", "YIELD", "DEPBAR",
                "MEMBAR", "ERRBAR", "CCTL", "ACQBULK", "BMOV", "S2R", "S2UR", "CS2R", "R2UR", "R2P", "P2R", "UR2UP"):
        return "control"
    if base in ALU:
        return "alu"
    return "other"


def parse(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        if cur is None:
            continue
        m = INSTR.match(line)
        if not m:
            continue
        addr, text = int(m.group(1), 16), m.group(2).strip()
        text = re.sub(r"^@!?U?P\d+\s+", "", text)          # predicate
        op = text.split()[0]
        tgt = None
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)\s*$", text)
            if t:
                tgt = int(t.group(1), 16)
        funcs[cur].append((addr, op, tgt))
    return funcs


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:   # noqa: BLE001
        return {n: n for n in names}


def census(instrs):
    """the innermost loop (backward branch) with the most MUFU instructions"""
    addr_index = {a: i for i, (a, _, _) in enumerate(instrs)}
    loops = []
    for i, (a, op, tgt) in enumerate(instrs):
        if tgt is not None and tgt <= a and tgt in addr_index:
            loops.append((addr_index[tgt], i))
    # innermost loops only (no other backward branch inside), the one with the most MUFU
    best = None
    for lo, hi in loops:
        if any(lo <= l2 and h2 <= hi and (l2, h2) != (lo, hi) for l2, h2 in loops):
            continue
        n_mufu = sum(1 for _, op, _ in instrs[lo:hi + 1] if op.startswith("MUFU"))
        key = (n_mufu, -(hi - lo))
        if n_mufu and (best is None or key > best[0]):
            best = (key, lo, hi)
    if best is None:
        return None
    _, lo, hi = best
    body = instrs[lo:hi + 1]
    counts = {}
    mufu_kinds = {}
    for _, op, _ in body:
        c = classify(op)
        counts[c] = counts.get(c, 0) + 1
        if c == "mufu":
            mufu_kinds[op] = mufu_kinds.get(op, 0) + 1
    total = len(body)
    packed = counts.get("packed_fp32x2", 0)
    # every loop that holds MUFU work, outermost last: the maturity loop of the curve kernels contains the five-pair group
    # twice (a straight-line copy for the default save stride, HW1F_FIXED_HALF, and the general group loop counted above)
    # plus the save point (5 shuffles, the polynomial, the exponential fall-back)
    nest = []
    for l2, h2 in sorted(loops, key=lambda x: x[1] - x[0]):
        b2 = instrs[l2:h2 + 1]
        n_m = sum(1 for _, op, _ in b2 if op.startswith("MUFU"))
        if n_m:
            nest.append({"range": [hex(instrs[l2][0]), hex(instrs[h2][0])], "instructions": len(b2), "mufu": n_m,
                         "packed_fp32x2": sum(1 for _, op, _ in b2 if op.split(".")[0] in PACKED),
                         "shuffles": sum(1 for _, op, _ in b2 if op.startswith("SHFL"))})
    return {"loop_address_range": [hex(instrs[lo][0]), hex(instrs[hi][0])], "instructions": total, "by_class": counts,
            "mufu_kinds": mufu_kinds, "dispatch_cycles": total + packed, "xu_cycles": 8 * counts.get("mufu", 0),
            "loops_with_mufu": nest}


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    funcs = parse(lib)
    names = demangle(list(funcs))
    out = {"library": os.path.relpath(lib, ROOT), "how": "cuobjdump -sass; loop = innermost backward branch with the most MUFU; "
           "dispatch cycles = instructions + packed FP32x2 instructions; XU cycles = 8 per MUFU (per warp, per SM sub-partition)",
           "kernels": {}}
    for mangled, instrs in funcs.items():
        name = names[mangled]
        if not any(k in name for k in ("fast_kernel", "bond_curve_kernel", "zbc_kernel", "pathwise_kernel", "fused_kernel")):
            continue
        c = census(instrs)
        if c is None:
            continue
        short = re.sub(r"\(int\)|\(bool\)", "", name)
        short = re.sub(r"\(.*", "", short).replace("void hw1f::", "")
        c["total_instructions_in_kernel"] = len(instrs)
        if "fast_kernel" in short or short.startswith("bond_curve") or short.startswith("zbc_kernel"):
            # one body = 5 Box-Muller pairs x 2 lanes x 2 antithetic twins = 40 path-steps (pathwise: 20 path-steps x 2 processes)
            c["per_path_step"] = {"dispatch_cycles": c["dispatch_cycles"] / 40.0, "mufu": c["by_class"].get("mufu", 0) / 40.0}
        out["kernels"][short] = c
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
