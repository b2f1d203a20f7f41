"""SASS opcode histogram per kernel of a cubin/.so (developer helper): python tools/sass_hist.py lib.so [filter]"""
import collections
import re
import subprocess
import sys

out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
flt = sys.argv[2] if len(sys.argv) > 2 else ""
cur, hist = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(2).split(".")[0]] += 1
for k, h in hist.items():
    if flt in k:
        print(k, sum(h.values()), dict(h.most_common(16)))
