#!/usr/bin/env python3
"""Per-function SASS statistics of an object / cubin (cuobjdump -sass output on stdin or as argv[1]): instruction count, code bytes,
local-memory traffic, IMAD / CALL counts.  Used for the profiles/r2_sass_*.txt summaries.  usage: tools/sass_stats.py file.sass [filter]"""
import re
import sys
from collections import Counter

txt = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for f in re.split(r"\n\s+Function : ", txt)[1:]:
    name = f.split("\n")[0]
    if flt and not re.search(flt, name):
        continue
    ins = re.findall(r"/\*([0-9a-f]{4,6})\*/\s+(?:@!?U?P\d\s+)?(\S+)", f)
    c = Counter(op.split(".")[0] for _, op in ins)
    wide = sum(1 for _, op in ins if op.startswith("IMAD.WIDE"))
    print(f"{name[:110]}\n    instructions {len(ins)}  code {len(ins) * 16} B  IMAD.WIDE {wide}  IMAD(other) {c['IMAD'] - wide}  IADD3 {c['IADD3']}  "
          f"STL {c['STL']}  LDL {c['LDL']}  CALL {c['CALL']}  MOV {c['MOV']}  LDG {c['LDG']}  STG {c['STG']}  BRA {c['BRA']}")
