"""Opcode census per function of a cubin / .so / .o: python scripts/sass_census.py <file> [name filter]"""
import collections
import re
import subprocess
import sys

txt = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for p in re.split(r"\n\s+Function : ", txt)[1:]:
    name = p.split("\n", 1)[0]
    if flt not in name:
        continue
    ops = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", p):
        ops[m.group(1)] += 1
    print(name, sum(ops.values()))
    print("  ", ops.most_common(18))
