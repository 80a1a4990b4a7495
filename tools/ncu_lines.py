"""Per-source-line instruction counts and stall samples of one profiled kernel.
Joins the SASS page of an ncu report (`ncu -i X.ncu-rep --page source --csv --print-source sass`)
with the line table of the same cubin (`nvdisasm -g`, extracted with `cuobjdump -xelf all`).
Usage: python tools/ncu_lines.py sass_page.csv nvdisasm_function.txt [top]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ia, iex, ismp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(data[0][ia], 16)
ex = {int(r[ia], 16) - base: (int(r[iex]), int(r[ismp]), r[hdr.index("Source")]) for r in data}
cur = None
per = defaultdict(lambda: [0, 0, 0, 0])
for line in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", line)
    if m and cur:
        a = int(m.group(1), 16)
        if a in ex:
            e = ex[a]
            p = per[cur]
            p[0] += e[0]; p[1] += e[1]; p[2] += 1
            if re.search(r"\b(DFMA|DMUL|DADD|DSETP)", e[2]):
                p[3] += e[0]
tot = sum(p[0] for p in per.values()); ts = sum(p[1] for p in per.values())
top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
src = {}
print("total warp instructions %d, samples %d" % (tot, ts))
for (k, p) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
    if k[0] not in src:
        try:
            src[k[0]] = open("pyrate_b200/csrc/" + k[0]).read().split("\n")
        except OSError:
            src[k[0]] = []
    text = src[k[0]][k[1] - 1].strip()[:90] if 0 < k[1] <= len(src[k[0]]) else ""
    print("%5.2f%% inst (fp64 %5.2f%%) %5.2f%% samples  %3d sass  %s:%d  %s" %
          (100.0 * p[0] / tot, 100.0 * p[3] / tot, 100.0 * p[1] / ts, p[2], k[0], k[1], text))
