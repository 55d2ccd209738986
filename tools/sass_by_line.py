#!/usr/bin/env python
"""Attribute the SASS instructions of one kernel's address range to source lines (no GPU needed):
    cuobjdump -xelf all pumi-pic_b200/_obj/pp_search.cu.o          # -> pp_search.sm_100a.cubin
    nvdisasm -g -c pp_search.sm_100a.cubin > all.txt
    python tools/sass_by_line.py all.txt k_walk_scsILi3ELi2ELb0ELi4 0x1f60 0x44e0 [0x3540 0x4460]
prints instruction counts per source line inside [lo, hi], the optional inner range listed apart."""
import collections
import re
import sys


def main():
    path, kernel, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
    ilo, ihi = (int(sys.argv[5], 16), int(sys.argv[6], 16)) if len(sys.argv) > 6 else (1, 0)
    inst = re.compile(r'^\s+/\*([0-9a-f]{4})\*/\s+(.*?);')
    line = re.compile(r'//## File "([^"]+)", line (\d+)')
    on, cur = False, None
    outer, inner = collections.Counter(), collections.Counter()
    for l in open(path):
        if ".section" in l and ".text." in l:
            on = kernel in l
            continue
        if not on:
            continue
        m = line.search(l)
        if m:
            cur = "%s:%s" % (m.group(1).split("/")[-1], m.group(2))
            continue
        m = inst.match(l)
        if m:
            a = int(m.group(1), 16)
            if lo <= a <= hi:
                (inner if ilo <= a <= ihi else outer)[cur] += 1
    for name, c in (("outer", outer), ("inner", inner)):
        if c:
            print("== %s: %d instructions" % (name, sum(c.values())))
            for k, v in c.most_common(25):
                print("%5d  %s" % (v, k))


if __name__ == "__main__":
    main()
