#!/usr/bin/env python
"""Per-instruction stall attribution from an ncu report (SASS source page): which instructions of a
kernel's hot loop the warps wait ON.  A sample is attributed to the instruction a warp is about to issue,
so the wait belongs to that instruction's producers.
usage: ncu_stalls.py <report.ncu-rep> <kernel-substring> [top N]"""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, key = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None and r and r[0].startswith("0x"):
            cur["rows"].append(r)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
    blk = [b for b in blocks if key in b["name"]][-1]
    hdr = blk["hdr"]
    col = {h: i for i, h in enumerate(hdr)}
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    ins = blk["rows"]
    tot_s = sum(int(r[col["# Samples"]]) for r in ins)
    exe = [int(r[col["Instructions Executed"]]) for r in ins]
    hot = sorted(exe)[int(0.8 * len(exe))]     # the loop body is most of the code: 80th percentile of the counts
    loop = [r for r in ins if int(r[col["Instructions Executed"]]) >= 0.5 * hot]
    print(f"kernel {blk['name'][:80]}")
    print(f"instructions {len(ins)}, hot-loop instructions {len(loop)}, samples {tot_s}, "
          f"in hot loop {sum(int(r[col['# Samples']]) for r in loop)}")
    by_reason = collections.Counter()
    for r in loop:
        for h in reasons:
            by_reason[h] += int(r[col[h]])
    tot = sum(by_reason.values())
    print("stall reasons over the hot loop (share of samples):")
    print("  " + "  ".join(f"{k[6:]} {100.0 * v / tot:.1f}%" for k, v in by_reason.most_common(12)))
    # by opcode of the instruction waited at
    by_op = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for r in loop:
        op = re.sub(r"^@!?U?P\w+\s+", "", r[col["Source"]].strip()).split()[0].split(".")[0]
        e = by_op[op]
        e[0] += 1
        e[1] += int(r[col["# Samples"]])
        for h in reasons:
            e[2][h] += int(r[col[h]])
    print("by opcode of the stalled instruction: count, samples, samples per instruction, top reasons")
    for op, (n, s, rs) in sorted(by_op.items(), key=lambda kv: -kv[1][1])[:22]:
        tr = " ".join(f"{k[6:]}:{v}" for k, v in rs.most_common(3))
        print(f"  {op:10s} {n:5d} {s:8d} {s / n:8.1f}   {tr}")
    print(f"top {top} instructions by samples:")
    for r in sorted(loop, key=lambda r: -int(r[col["# Samples"]]))[:top]:
        rs = collections.Counter({h: int(r[col[h]]) for h in reasons})
        tr = " ".join(f"{k[6:]}:{v}" for k, v in rs.most_common(3) if v)
        print(f"  {r[0][-5:]} {int(r[col['# Samples']]):6d}  {r[col['Source']].strip()[:70]:70s} {tr}")


if __name__ == "__main__":
    main()
