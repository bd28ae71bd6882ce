#!/usr/bin/env python
"""Static code size of one kernel per source line (nvdisasm -g -c output): what fills the instruction cache.
  python profiles/codesize.py dis.txt <mangled-name-substring> [top]"""
import collections
import sys
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from attribute import line_table

tab = line_table(sys.argv[1], sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
agg = collections.Counter(v[0] for v in tab.values())
print(f"# {sys.argv[2]}: {len(tab)} SASS instructions, {len(tab) * 16 / 1024:.1f} KB")
for k, n in agg.most_common(top):
    print(f"  {k[0] + ':' + str(k[1]) if k else '?':34s} {n:6d}  {n * 16 / 1024:6.1f} KB")
