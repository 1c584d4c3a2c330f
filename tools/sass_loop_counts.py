#!/usr/bin/env python
"""Instruction counts of the innermost loops of a kernel, from its SASS (cuobjdump -sass).
   python tools/sass_loop_counts.py constant_ph_b200/csrc/pair.o 'pair_eval_kernelILi1ELi1ELb1E'
A loop is a backward branch; only loops that contain no other backward branch are listed.  fp64 = opcodes of the
fp64 pipe (DFMA DMUL DADD DSETP ...), the two MUFU.*64H seeds are listed separately."""
import re
import subprocess
import sys
from collections import Counter


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, rows = None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn is None or pat not in fn:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            rows.append((int(m.group(1), 16), m.group(2).strip()))
    if not rows:
        sys.exit("no function matches %r" % pat)
    loops = []
    for addr, ins in rows:
        m = re.search(r"\bBRA(?:\.U)?\b.*?0x([0-9a-f]+)", ins)
        if m and int(m.group(1), 16) <= addr:
            loops.append((int(m.group(1), 16), addr))
    inner = [l for l in loops if not any(o != l and l[0] <= o[0] and o[1] <= l[1] for o in loops)]
    print("%s: %d instructions, %d loops, %d innermost" % (pat, len(rows), len(loops), len(inner)))
    for lo, hi in inner:
        body = [ins for a, ins in rows if lo <= a <= hi]
        ops = Counter()
        for ins in body:
            op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]
            ops[op.split(".")[0] if not op.startswith("MUFU") else op] += 1
        fp64 = sum(v for k, v in ops.items() if re.match(r"D(FMA|MUL|ADD|SETP|MNMX)$", k))
        mufu = sum(v for k, v in ops.items() if k.startswith("MUFU"))
        print("  loop 0x%04x-0x%04x: %3d instructions, fp64 %3d, MUFU %d, LDG %d, LDS %d, other %d" % (
            lo, hi, len(body), fp64, mufu, ops["LDG"], ops["LDS"], len(body) - fp64 - mufu - ops["LDG"] - ops["LDS"]))


if __name__ == "__main__":
    main()
