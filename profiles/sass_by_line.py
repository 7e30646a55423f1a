#!/usr/bin/env python
"""Static SASS instruction counts per source line of one kernel (no GPU needed):
    cuobjdump -xelf all nspeech_b200/libnspeech_b200.so && nvdisasm --print-line-info *.cubin > all.sass
    python profiles/sass_by_line.py all.sass <mangled kernel name substring> [top N]
"""
import collections
import re
import sys

path, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cnt, ops, opc = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
on, cur, total = False, None, 0
for ln in open(path, errors="replace"):
    if ln.startswith(".text."):
        on = key in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
    if m and cur:
        op = m.group(2)
        cnt[cur] += 1
        total += 1
        ops[cur][op.split(".")[0]] += 1
        opc[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("LDS", "STS", "LDG", "STG", "MUFU", "UTC", "UBLK", "LDTM", "STTM", "SYNCS")) and "." in op else "")] += 1
print("static instructions:", total)
byfile = collections.Counter()
for (f, l), c in cnt.items():
    byfile[f] += c
print("by file:", dict(byfile))
print("by opcode:", dict(opc.most_common(40)))
for (f, l), c in sorted(cnt.items(), key=lambda x: -x[1])[:top]:
    print("%-18s %5d  %5d  %s" % (f, l, c, dict(ops[(f, l)].most_common(6))))
