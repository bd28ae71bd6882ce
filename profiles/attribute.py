#!/usr/bin/env python
"""Per-source-line attribution of one kernel: pairs the SASS rows of an `ncu --page source --csv` dump with
`nvdisasm -g -c` of the cubin the library was linked from (ncu's own source page needs the .cu on the box).

  cuobjdump -xelf all build/b200rt_api.cu.o ; nvdisasm -g -c b200rt_api.sm_100a.cubin > dis.txt
  ncu -i rep.ncu-rep --page source --csv --kernel-name k_shade7_b > src.csv
  python profiles/attribute.py dis.txt src.csv <mangled-name-substring> [top]
"""
import collections
import csv
import re
import sys


def line_table(dis, func):
    tab, cur, inside = {}, None, False
    for ln in open(dis):
        if ln.startswith(".text.") and ln.rstrip().endswith(":"):
            inside = func in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            if "inlined at" not in ln or cur is None:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(\S+)", ln)
        if m:
            tab[int(m.group(1), 16)] = (cur, m.group(2))
    return tab


def main():
    dis, src, func = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    tab = line_table(dis, func)
    rows = [r for r in csv.reader(open(src)) if r and r[0].startswith("0x")]
    hdr = next(r for r in csv.reader(open(src)) if r and r[0] == "Address")
    ia, isamp, iexec, ithr = (hdr.index(k) for k in ("Address", "# Samples", "Instructions Executed",
                                                      "Thread Instructions Executed"))
    base = int(rows[0][ia], 16)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    seen = set()
    for r in rows:
        off = int(r[ia], 16) - base
        if off in seen:          # a second launch of the same kernel in the dump
            break
        seen.add(off)
        key = tab.get(off, (None, "?"))[0]
        a = agg[key]
        a[0] += 1
        a[1] += int(r[isamp] or 0)
        a[2] += int(r[iexec] or 0)
        a[3] += int(r[ithr] or 0)
    tot = [sum(a[i] for a in agg.values()) for i in range(4)]
    print(f"# {func}: {tot[0]} SASS instructions ({tot[0] * 16 / 1024:.1f} KB), {tot[1]} samples, "
          f"{tot[2]} warp instructions, {tot[3] / max(1, tot[2]):.1f} active lanes/instruction")
    for title, idx in (("static code size", 0), ("stall samples", 1), ("executed warp instructions", 2)):
        print(f"## top source lines by {title}")
        for key, a in sorted(agg.items(), key=lambda kv: -kv[1][idx])[:top]:
            name = f"{key[0]}:{key[1]}" if key else "?"
            print(f"  {name:34s} sass {a[0]:6d} ({a[0] / tot[0] * 100:4.1f}%)  samples {a[1] / max(1, tot[1]) * 100:5.1f}%  "
                  f"executed {a[2] / max(1, tot[2]) * 100:5.1f}%  lanes {a[3] / max(1, a[2]):5.1f}")


if __name__ == "__main__":
    main()
