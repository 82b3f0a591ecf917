#!/usr/bin/env python
"""Static instruction mix of one kernel of libb200ks.so (cuobjdump -sass): counts per mnemonic.
    python profiles/sass_mix.py 'dslash_half_kernelILi0ELi0ELi7' [lib]
The 16-bit stencil is issue-bound, so its SASS length per site is the yardstick."""
import collections
import re
import subprocess
import sys

pat = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "milc_qcd_b200/libb200ks.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, counts = None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            counts[m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).startswith(("LDG", "STG", "LDL", "STL")) and "." in m.group(1) else "")] += 1
tot = sum(counts.values())
print("kernel pattern %s: %d instructions" % (pat, tot))
for k, v in counts.most_common(40):
    print("  %-14s %5d" % (k, v))
