"""SASS operand-bandwidth estimate of an inner loop (round 2, DESIGN.md sections 5-6).

    cuobjdump -sass -fun '<mangled kernel>' u_pol_b200/lib/libupol_b200.so > k.sass
    python scripts/sass_cost.py k.sass <first address hex> <last address hex> [fp64]

fp32 mode: every FFMA2/FMUL2/FADD2 (and scalar FFMA/FMUL/FADD) of the address range is charged
max(2.07 (1.07 scalar), 0.52 x register operand words) cycles per warp instruction per sub-partition - the fit to
profiles/r02_ubench2.txt (64-bit register pair = 2 words, R.F32 broadcast = 1, immediates 0, a register repeated
inside the instruction or flagged .reuse by the previous one counts once).  fp64 mode: DFMA/DMUL/DADD cost 2.22
cycles, 3.09 with three distinct register operands.  MUFU count (8 XU cycles each) and issue slots are printed
next to it.  An estimate for choosing between formulations without GPU time, not a measurement."""
import re
import sys


def rows_of(path, a, b):
    out = []
    for l in open(path):
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?)\s*(/\*.*)?$", l)
        if not m:
            continue
        x = int(m.group(1), 16)
        if a <= x <= b:
            out.append(re.sub(r"^@!?U?P\d+\s+", "", m.group(2)).rstrip(" ;"))
    return out


def fp32_cost(rows):
    tot, n, prev = 0.0, 0, {}
    for t in rows:
        op = t.split()[0]
        base = op.split(".")[0]
        if base not in ("FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD"):
            continue
        srcs = [o.strip() for o in t[len(op):].split(",")[1:]]
        words, cur, seen = 0, {}, set()
        for slot, o in enumerate(srcs):
            m = re.match(r"^[-|]*(R\d+)(\.reuse)?((?:\.F32x2\.HI_LO)|(?:\.F32))?", o)
            if not m:
                continue
            reg, reuse, kind = m.groups()
            w = 2 if kind == ".F32x2.HI_LO" else 1
            if prev.get(slot) == reg or (reg, kind) in seen:
                w = 0
            seen.add((reg, kind))
            words += w
            if reuse:
                cur[slot] = reg
        prev = cur
        tot += max(2.07 if base.endswith("2") else 1.07, 0.52 * words)
        n += 1
    return n, tot


def fp64_cost(rows):
    tot, n, n3, prev = 0.0, 0, 0, {}
    for t in rows:
        op = t.split()[0]
        if op.split(".")[0] not in ("DFMA", "DMUL", "DADD"):
            continue
        regs, cur = set(), {}
        for slot, o in enumerate(t[len(op):].split(",")[1:]):
            m = re.match(r"^\s*[-|]*(R\d+)(\.reuse)?", o)
            if not m:
                continue
            if prev.get(slot) != m.group(1):
                regs.add(m.group(1))
            if m.group(2):
                cur[slot] = m.group(1)
        prev = cur
        n += 1
        n3 += len(regs) >= 3
        tot += 3.09 if len(regs) >= 3 else 2.22
    return n, n3, tot


if __name__ == "__main__":
    rows = rows_of(sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16))
    mufu = sum("MUFU" in r for r in rows)
    if len(sys.argv) > 4 and sys.argv[4] == "fp64":
        n, n3, tot = fp64_cost(rows)
        print(f"{len(rows)} issue slots, {n} DP instructions ({n3} with three distinct registers): ~{tot:.0f} cycles, floor {n * 2.22:.0f}")
    else:
        n, tot = fp32_cost(rows)
        print(f"{len(rows)} issue slots, {n} FMA-pipe instructions: ~{tot:.0f} cycles; {mufu} MUFU = {8 * mufu} XU cycles")
