#!/usr/bin/env python3
"""Static opcode histogram of one kernel in a .so/.cubin:  python tools/sass_static.py <lib> <substring of mangled name> [--dump]"""
import collections, re, subprocess, sys
lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, errors="replace").stdout
cur, hist, lines = None, collections.Counter(), []
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(2).split(".")[0]
            hist[op] += 1
            lines.append(ln.rstrip()[:110])
print(sum(hist.values()), "instructions")
for k, v in hist.most_common(40):
    print(f"  {k:10s} {v}")
if "--dump" in sys.argv:
    print("\n".join(lines))
if "--loops" in sys.argv:
    addr = []
    for ln in lines:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
        addr.append((int(m.group(1), 16), m.group(2)))
    for a, txt in addr:
        m = re.search(r"BRA.*?0x([0-9a-f]+)", txt)
        if m and int(m.group(1), 16) < a:
            t = int(m.group(1), 16)
            body = [x for x in addr if t <= x[0] <= a]
            h = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", x[1]).split()[0].split(".")[0] for x in body)
            print(f"loop 0x{t:x}..0x{a:x}: {len(body)} instr:", dict(h.most_common(14)))
