#!/usr/bin/env python
"""Instruction mix of one kernel in an object file, normalised per DMUL-group.

    python tools/sass_mix.py libprosic_b200/_build/pairhmm_inst_01.o 'pairhmm_kernelILi5ELb0ELb1ELb0E' [per]

Prints opcode counts of the whole function and, if `per` is given, counts divided by
(count of DMUL / per) -- e.g. per=5 DMUL per cell for the no-extension pair-HMM kernels.
"""
import re
import subprocess
import sys
from collections import Counter


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    per = float(sys.argv[3]) if len(sys.argv) > 3 else 0
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, ops = None, Counter()
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or pat not in cur:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m:
            continue
        ins = m.group(1).split()
        op = ins[1] if ins[0].startswith("@") else ins[0]
        ops[op.split(".")[0]] += 1
    tot = sum(ops.values())
    cells = ops["DMUL"] / per if per else 0
    print("total %d instructions%s" % (tot, ", %.0f cells, %.1f instr/cell" % (cells, tot / cells) if cells else ""))
    for op, n in ops.most_common(24):
        print("  %-10s %6d %s" % (op, n, "%.2f" % (n / cells) if cells else ""))


if __name__ == "__main__":
    main()
