#!/usr/bin/env python
"""Dynamic opcode histogram of one kernel from `ncu -i rep --page source --csv --kernel-name regex:K` (stdin or file):
warp-instructions executed per opcode (IMAD split by flavour), per "unit" (pass --units N, e.g. warps launched)."""
import collections
import csv
import re
import sys


def main():
    args = sys.argv[1:]
    units = 1.0
    if "--units" in args:
        i = args.index("--units")
        units = float(args[i + 1])
        del args[i:i + 2]
    rows = list(csv.reader(open(args[0]) if args else sys.stdin))
    h = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[h]
    ia, isrc, iss = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    cnt, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in rows[h + 1:]:
        if len(r) <= ia or not r[ia].isdigit():
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+)*)", r[isrc].strip())
        if not m:
            continue
        parts = m.group(2).split(".")
        key = parts[0]
        if key == "IMAD":
            for f in ("MOV", "SHL", "IADD", "WIDE", "HI"):
                if f in parts:
                    key = "IMAD." + f
                    break
        n = int(r[ia])
        cnt[key] += n
        samp[key] += int(r[iss] or 0)
        tot += n
    print("total warp-instructions %d  (%.1f per unit)" % (tot, tot / units))
    for k, v in cnt.most_common(45):
        print("%-12s %10.1f per unit  %5.1f%%  stall samples %d" % (k, v / units, 100.0 * v / tot, samp[k]))


if __name__ == "__main__":
    main()
