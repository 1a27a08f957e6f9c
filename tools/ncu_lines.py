#!/usr/bin/env python
"""Join an ncu SASS source page (--page source --csv) with nvdisasm line info so
that stall samples can be read per CUDA source line.
usage: ncu_lines.py <report.ncu-rep> <cubin> <mangled-kernel-substring> [top]"""
import csv, re, subprocess, sys, collections
rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# walk the function: remember the current //## File ... line N
lines = []; cur = None; infun = False
for l in txt.splitlines():
    if l.startswith("\t.text.") or l.startswith(".text."):
        infun = kname in l
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)))
    if infun and re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        lines.append((cur, l.strip()))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]; idx = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr)]
print(f"sass rows {len(data)}, nvdisasm instrs {len(lines)}", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = 0
for i, r in enumerate(data):
    s = int(r[idx["# Samples"]] or 0); ins = int(r[idx["Instructions Executed"]] or 0)
    key = lines[i][0] if i < len(lines) else None
    a = agg[key]; a[0] += s; a[1] += ins; tot += s
    for c in stall_cols:
        v = int(r[idx[c]] or 0)
        if v: a[2][c] += v
src = {}
for key, (s, ins, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if key and key[0] not in src:
        try: src[key[0]] = open(f"/root/repo/transhuman_b200/csrc/{key[0]}").read().splitlines()
        except Exception: src[key[0]] = []
    text = src[key[0]][key[1] - 1].strip()[:90] if key and src[key[0]] and key[1] <= len(src[key[0]]) else ""
    tops = ",".join(f"{k[6:]}:{v}" for k, v in st.most_common(3))
    print(f"{s:7d} {100 * s / max(tot, 1):5.1f}% inst {ins:9d} {str(key):28s} {tops:40s} | {text}")
print("total samples", tot)
