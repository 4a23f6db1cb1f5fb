#!/usr/bin/env python3
"""Per-phase maximum / minimum over the ranks of the LAST proof in a ZKAES_TRACE=1 log of a multi-rank run (development aid).
    trace_summary.py <raw stderr log> <ranks>"""
import collections
import re
import sys


def main():
    path, n = sys.argv[1], int(sys.argv[2])
    rows = collections.OrderedDict()
    for line in open(path):
        m = re.match(r"\[zkaes\] (.{28})\s+([0-9.]+) ms", line)
        if m and not m.group(1).startswith("keys"):
            rows.setdefault(m.group(1).strip(), []).append(float(m.group(2)))
    total = 0.0
    for k, v in rows.items():
        last = v[-n:]
        total += max(last)
        print(f"{k:28s} {max(last):9.2f} ms max over ranks   {min(last):9.2f} min")
    print(f"{'sum of the maxima':28s} {total:9.2f} ms")


if __name__ == "__main__":
    main()
