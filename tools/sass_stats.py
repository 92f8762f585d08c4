#!/usr/bin/env python
"""Static SASS statistics of a kernel's hot loop (no GPU needed): instructions per marched row,
per-pipe counts and the sum of the stall fields (= cycles one warp alone needs per row, the
inverse of the ILP ptxas found).  usage: sass_stats.py <lib.so> <mangled-substring> [rows-per-loop]"""
import collections
import re
import subprocess
import sys


def parse(lib, key):
    names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
    ins, cur, regs = [], None, {}
    i = 0
    while i < len(names):
        ln = names[i]
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            cur = m.group(1)
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", ln)
        if m and cur and key in cur and i + 1 < len(names):
            m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", names[i + 1])
            if m2:
                ins.append((int(m.group(1), 16), m.group(2).strip(), int(m2.group(1), 16)))
                i += 2
                continue
        i += 1
    return ins


def main():
    lib, key = sys.argv[1], sys.argv[2]
    rows = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    ins = parse(lib, key)
    addr = {a: n for n, (a, _, _) in enumerate(ins)}
    # the hot loop = the longest backward branch
    best = None
    for n, (a, op, hi) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", op)
        if m:
            t = int(m.group(1), 16)
            # the first large loop: the fast pass (the SAFE redo pass is laid out after it)
            if t < a and t in addr and a - t > 16 * 200 and best is None:
                best = (t, a)
    lo, hi_ = best
    loop = [x for x in ins if lo <= x[0] <= hi_]
    cnt = collections.Counter()
    stall = 0
    for a, op, hi in loop:
        o = re.sub(r"^@!?U?P\w+\s+", "", op).split()[0].split(".")[0]
        cnt[o] += 1
        stall += (hi >> 41) & 0xF
    n = len(loop)
    dp = sum(cnt[k] for k in ("DADD", "DMUL", "DFMA", "DSETP"))
    alu = sum(cnt[k] for k in ("FSEL", "SEL", "LOP3", "ISETP", "IADD3", "PLOP3", "FSETP", "SHF", "LEA", "VIADD",
                               "PRMT", "FMNMX", "IABS", "R2P", "P2R", "MOV", "VIMNMX"))
    print(f"loop 0x{lo:x}..0x{hi_:x}: {n} instr = {n/rows:.1f}/row; stall-sum {stall/rows:.0f}/row "
          f"({stall/n:.2f}/instr); DP {dp/rows:.1f} ALU {alu/rows:.1f}/row")
    print("  " + " ".join(f"{k}:{v/rows:.1f}" for k, v in cnt.most_common(24)))


if __name__ == "__main__":
    main()
