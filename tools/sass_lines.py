#!/usr/bin/env python3
"""Executed warp instructions per SOURCE line of one kernel.

ncu keeps per-SASS-instruction counters in a `--set full` report; `nvdisasm -g` prints the line table of the same cubin.
Both list the kernel's instructions in address order, so they join by position:

  cuobjdump -xelf all cavif_rs_b200/libcavif_b200.so            # -> filters.sm_100a.cubin, rdo.sm_100a.cubin, ...
  nvdisasm -g filters.sm_100a.cubin > filters_lines.txt
  ncu -i report.ncu-rep --page source --csv --kernel-name regex:dbcdef > dbcdef.csv
  tools/sass_lines.py filters_lines.txt _Z18b200_dbcdef_kernelIhEvPK10B200RdoJobPKhi dbcdef.csv [top N]

The library must be the build the report was captured from (compiled with -lineinfo, as the Makefile does)."""
import collections
import csv
import re
import sys


def main():
    lines_file, func, ncu_csv = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    insts, cur, infunc = [], None, False
    for l in open(lines_file):
        if l.startswith(".text."):
            infunc = l.strip().rstrip(":") == ".text." + func
            continue
        if not infunc:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            insts.append((m.group(2).strip(), cur))
    rows = list(csv.reader(open(ncu_csv)))
    hdr = rows[1]
    isrc, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = [(r[isrc].strip(), int(r[ie]), int(r[isamp])) for r in rows[2:] if len(r) > ie and r[ie].isdigit()][:len(insts)]
    assert len(data) == len(insts), "report and cubin disagree on the instruction count: not the same build"
    mism = sum(1 for (t, _), (s, _, _) in zip(insts, data) if t.split()[-1][:3] != s.split()[-1][:3] and t.split()[0][:3] != s.split()[0][:3])
    tot, tots = sum(d[1] for d in data), sum(d[2] for d in data)
    by_line, by_file = collections.defaultdict(lambda: [0, 0, 0]), collections.defaultdict(lambda: [0, 0])
    for (_, ln), (_, e, sm) in zip(insts, data):
        a = by_line[ln]; a[0] += e; a[1] += sm; a[2] += 1
        b = by_file[ln[0] if ln else "?"]; b[0] += e; b[1] += sm
    print("kernel %s: %d SASS instructions, %d executed warp instructions, %d stall samples (opcode mismatches: %d)" % (func, len(insts), tot, tots, mism))
    for f, a in sorted(by_file.items(), key=lambda x: -x[1][0]):
        print("  %-32s inst %5.1f %%  samples %5.1f %%" % (f, a[0] * 100.0 / tot, a[1] * 100.0 / max(tots, 1)))
    print("top %d source lines by executed instructions" % top)
    for ln, a in sorted(by_line.items(), key=lambda x: -x[1][0])[:top]:
        print("  %-32s inst %5.1f %%  samples %5.1f %%  (%d SASS)" % ("%s:%d" % ln if ln else "?", a[0] * 100.0 / tot, a[1] * 100.0 / max(tots, 1), a[2]))


if __name__ == "__main__":
    main()
